"""Helper of test_gpu_variants.py: one cascade (N = 4096 unless given) against the CPU oracle, run in a fresh process so that
the library reads the SSPSD_* environment switches at handle creation."""
import sys

import numpy as np

from conftest import uniform_noise
from oracle import binding as orc
import stabilizer_stream_b200 as sp

det = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
x = uniform_noise(300 * n + 123, 31) + np.float32(0.1)
g = sp.PsdCascade(n)
g.set_detrend(sp.Detrend(det))
o = orc.Cascade(n, 1)
o.set_detrend(det)
for a, b in ((0, 100_001), (100_001, 700_000), (700_000, x.size)):
    g.process(x[a:b])
    o.process(x[a:b])
p, br = g.psd(sp.MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))
po, bo = o.psd(True, 0, True)
assert [k.count for k in br] == [k.count for k in bo]
worst = 0.0
for k in br:
    if k.count:
        sl = slice(k.start, k.start + len(k.bins))
        w = po[sl].astype(np.float64)
        rel = (np.abs(p[sl] - w) - 1e-5 * np.median(w)) / w
        worst = max(worst, float(np.max(rel[4:])))
assert worst < 1e-4, worst
print("variant ok", worst)
