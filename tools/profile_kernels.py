#!/usr/bin/env python3
"""Small driver for ncu captures of the stage-0 kernels (run under gpurun, see profiles/README.md): processes a few
2^26-sample batches through default cascades of the given FFT sizes.  The first PSD / decimator launches of each
process() call are the full-size stage-0 launches.

    python tools/profile_kernels.py [n_fft[,n_fft...]] [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from stabilizer_stream_b200 import MergeOpts, PsdCascade  # noqa: E402

sizes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4096").split(",")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = (torch.rand(1 << 26, device="cuda") - 0.5) * (12 ** 0.5)
for n in sizes:
    c = PsdCascade(n, deep_defer=1)
    for _ in range(reps):
        c.process(x)
    p, b = c.psd(MergeOpts())
    print("n", n, "stages", [(k.decimation, k.count) for k in b], "bins", p.size)
