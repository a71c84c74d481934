"""The library's environment switches select other kernels / stream layouts for the same result: the tiled
radix-8 and radix-16 PSD kernels at N = 4096 (SSPSD_K2), everything on one stream (SSPSD_OVERLAP=0), no separate
PSD stream (SSPSD_PSD_STREAM=0).  Each variant must meet the same parity bound as the default."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("env,det", [({"SSPSD_K2": "r16"}, 0), ({"SSPSD_K2": "r16"}, 3), ({"SSPSD_K2": "r8"}, 0),
                                     ({"SSPSD_OVERLAP": "0"}, 1), ({"SSPSD_PSD_STREAM": "0"}, 2), ({"SSPSD_OVERLAP": "2"}, 0)])
def test_kernel_and_stream_variants(env, det):
    e = dict(os.environ)
    e.update(env)
    e["PYTHONPATH"] = os.pathsep.join([os.path.dirname(HERE), HERE, e.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, os.path.join(HERE, "variant_check.py"), str(det)], env=e, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout[-500:] + r.stderr[-1500:]
