#!/usr/bin/env python3
"""Smallest run that touches every kernel family once, for `compute-sanitizer --tool memcheck` (one tool per call):
cascades at N = 512 (warp-level kernel) and N = 4096 (ring kernel) with ragged host / device feeding, the persistent
decimator, a deterministic-mode handle, frame decode, a time-chunked group of three ranks on one GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import stabilizer_stream_b200 as sp  # noqa: E402
from frames_util import make_frames  # noqa: E402

rng = np.random.default_rng(1)
for n in (512, 4096, 1024):
    x = ((rng.random(300 * n + 17, dtype=np.float32) - 0.5) * 3.4641).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    for det, kw in ((0, {}), (3, {"deterministic": True}), (2, {"deep_defer": 1})):
        c = sp.PsdCascade(n, **kw)
        c.set_detrend(sp.Detrend(det))
        c.process(x[:5 * n + 3])
        c.process(xd[5 * n + 3:200 * n])
        c.process(x[200 * n:])
        p, b = c.psd()
        assert np.all(np.isfinite(p)) and len(b) >= 2
data, flen, stride, _ = make_frames(1, 22, 300, seed=5, drop_every=11)
loss = sp.Loss()
fmt, traces, ok = sp.FrameDecoder().decode(data, flen, loss)
assert ok == 300
g = sp.Group(512, devices=[0, 0, 0], mode=sp.ShardMode.TIME)
g.time_plan(1_500_000)
g.time_process_noise()
g.time_finish()
p, b = g.psd(0)
assert np.all(np.isfinite(p))
torch.cuda.synchronize()
print("sanitize smoke ok")
