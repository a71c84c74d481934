"""The library's environment switches select other kernels / stream layouts for the same result: the tiled
radix-8 and radix-16 PSD kernels (SSPSD_K2; at N = 512 "r8" is the generic kernel the warp-level one replaced), the
decimator generations (SSPSD_K3: tiled, TMA-staged persistent with 960 / 640 outputs per tile, cp.async scatter),
deferral of the deep stages off / tiny (SSPSD_DEFER), everything on one stream (SSPSD_OVERLAP=0), no separate PSD
stream (SSPSD_PSD_STREAM=0).  Each variant must meet the same parity bound as the default."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("env,det,n", [({"SSPSD_K2": "r16"}, 0, 4096), ({"SSPSD_K2": "r16"}, 3, 4096), ({"SSPSD_K2": "r8"}, 0, 4096),
                                       ({"SSPSD_OVERLAP": "0"}, 1, 4096), ({"SSPSD_PSD_STREAM": "0"}, 2, 4096),
                                       ({"SSPSD_OVERLAP": "2"}, 0, 4096),
                                       ({"SSPSD_K3": "tiled"}, 0, 4096), ({"SSPSD_K3": "tma640"}, 3, 4096),
                                       ({"SSPSD_K3": "async960"}, 1, 4096), ({"SSPSD_K3": "async640"}, 0, 512),
                                       ({"SSPSD_K3": "pf960"}, 2, 4096), ({"SSPSD_K3": "tma768"}, 1, 4096),
                                       ({"SSPSD_CARRY_KERNEL": "1"}, 3, 4096), ({"SSPSD_CARRY_KERNEL": "1"}, 0, 512), ({"SSPSD_K3": "pf640"}, 3, 512),
                                       ({"SSPSD_DEFER": "1"}, 2, 4096), ({"SSPSD_DEFER": "4096"}, 0, 512),
                                       ({"SSPSD_K2": "ring1"}, 3, 4096), ({"SSPSD_K2": "ring1", "SSPSD_OVERLAP": "0"}, 1, 4096),
                                       ({"SSPSD_K2": "ring1x5"}, 2, 4096),
                                       ({"SSPSD_K2": "r8"}, 3, 512), ({"SSPSD_K2": "ring", "SSPSD_OVERLAP": "0"}, 2, 512)])
def test_kernel_and_stream_variants(env, det, n):
    e = dict(os.environ)
    e.update(env)
    e["PYTHONPATH"] = os.pathsep.join([os.path.dirname(HERE), HERE, e.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, os.path.join(HERE, "variant_check.py"), str(det), str(n)], env=e, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout[-500:] + r.stderr[-1500:]
