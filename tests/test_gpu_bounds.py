"""Memory safety without compute-sanitizer (closed on this pool: profiles/r02_compute_sanitizer_refused.log): the
bounds-checking build `libsspsd_bounds.so` (-DSSPSD_BOUNDS) asserts on the DEVICE every global index a kernel forms
from run-time bookkeeping -- stream reads (carry / fresh extents, TMA bulk-copy ranges), decimator outputs, carry
copies, decoded traces.  A failing assert poisons the CUDA context, so the run below fails loudly.  It drives every
kernel family through ragged host / device feeding, the decimator generations, deterministic and deferred modes,
frame decode and a time-chunked group (tools/sanitize_smoke.py), plus seeded API walks from the fuzz suite."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BOUNDS = os.path.join(ROOT, "stabilizer_stream_b200", "libsspsd_bounds.so")


def _env(extra=None):
    e = dict(os.environ, SSPSD_LIB=BOUNDS)
    e["PYTHONPATH"] = os.pathsep.join([ROOT, os.path.join(ROOT, "tests"), e.get("PYTHONPATH", "")])
    e.update(extra or {})
    return e


@pytest.mark.parametrize("env", [{}, {"SSPSD_K3": "tiled"}, {"SSPSD_K3": "async640"}, {"SSPSD_K3": "tma640", "SSPSD_K2": "r8"},
                                 {"SSPSD_K2": "r16", "SSPSD_OVERLAP": "0"}, {"SSPSD_K3": "pf960", "SSPSD_K2": "ring1"},
                                 {"SSPSD_K3": "pf640", "SSPSD_K2": "ring1x5"}])
def test_bounds_build_smoke(env):
    if not os.path.exists(BOUNDS):
        pytest.skip("libsspsd_bounds.so not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_smoke.py")], env=_env(env), capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "sanitize smoke ok" in r.stdout, r.stdout[-500:] + r.stderr[-2000:]


def test_bounds_build_fuzz_walks():
    if not os.path.exists(BOUNDS):
        pytest.skip("libsspsd_bounds.so not built")
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", os.path.join(ROOT, "tests", "test_gpu_fuzz.py"),
                        os.path.join(ROOT, "tests", "test_gpu_timechunk.py"), os.path.join(ROOT, "tests", "test_gpu_group.py"),
                        os.path.join(ROOT, "tests", "test_gpu_decode.py")],
                       env=_env({"SSPSD_FUZZ_EXTRA": "24"}), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
