#!/usr/bin/env python3
"""Debug aid for tests/test_gpu_fuzz.py: replays one seeded API walk, prints every operation, and at each
readout reports the worst bin per stage against the CPU oracle.  usage: fuzz_trace.py SEED N [-q]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from conftest import uniform_noise  # noqa: E402
from oracle import binding as oracle  # noqa: E402
import stabilizer_stream_b200 as sp  # noqa: E402

seed, n = int(sys.argv[1]), int(sys.argv[2])
quiet = "-q" in sys.argv
rng = np.random.default_rng(seed)
x = uniform_noise(260 * n * 8, 100 + seed) + np.float32(0.05)
xd = torch.from_numpy(x).cuda()
mb = int(rng.integers(20 * n, 400 * n))
print("max_batch", mb, flush=True)
g = sp.PsdCascade(n, hbf=sp.Hbf(seed & 1), host_stage=1 << 13, max_batch=mb)
o = oracle.Cascade(n, seed & 1)
pos = 0
step = 0
det = 0
avg = (2 ** 32 - 1, 2 ** 32 - 1)


def log(*a):
    if not quiet:
        print(*a, flush=True)


while pos < x.size:
    step += 1
    r = rng.random()
    if r < 0.55:
        k = int(rng.integers(0, 30 * n)) if rng.random() < 0.8 else int(rng.integers(0, 9))
        dev = rng.random() >= 0.5
        log(step, "process", "dev" if dev else "host", k, "pos", pos)
        g.process(xd[pos:pos + k] if dev else x[pos:pos + k])
        o.process(x[pos:pos + k])
        pos += k
    elif r < 0.65:
        det = int(rng.integers(0, 4))
        log(step, "detrend", det)
        g.set_detrend(sp.Detrend(det))
        o.set_detrend(det)
    elif r < 0.72:
        avg = (int(rng.choice([2 ** 32 - 1, 3, 17, 200, 0])), int(rng.choice([2 ** 32 - 1, 2 ** 32 - 2, 5000, 64])))
        log(step, "avg", avg)
        g.set_avg(sp.AvgOpts(limit=avg[0], count=avg[1]))
        o.set_avg(*avg)
    elif r < 0.84:
        p, b = g.psd(sp.MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))
        po, bo = o.psd(True, 0, True)
        for k in b:
            if not k.count:
                continue
            sl = slice(k.start, k.start + len(k.bins))
            w = po[sl].astype(np.float64)
            rel = (np.abs(p[sl] - w) - 1e-5 * np.median(w)) / np.maximum(w, 1e-300)
            i = int(np.argmax(rel[4:])) + 4
            flag = "  <-- FAIL" if rel[i] > 1e-4 else ""
            if flag or not quiet:
                print(step, "psd dec=%d count=%d avg=%d det=%d: worst bin %d rel %.3g got %.6g want %.6g median %.4g%s"
                      % (k.decimation, k.count, k.avg, det, i, rel[i], p[sl][i], w[i], np.median(w), flag), flush=True)
    elif r < 0.90:
        log(step, "clone")
        g = g.clone()
        o = o.clone()
    elif r < 0.93 and pos > 0:
        log(step, "reset")
        g.reset()
        o = oracle.Cascade(n, seed & 1)
        o.set_detrend(0)
        det = 0
        g.set_detrend(sp.Detrend(0))
        g.set_avg(sp.AvgOpts())
print("done")
