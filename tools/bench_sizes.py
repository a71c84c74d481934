#!/usr/bin/env python3
"""Device-resident cascade throughput for every supported FFT size (200e6-sample stream, default
options).  Prints one line per size.  Usage: python tools/bench_sizes.py [sizes...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from stabilizer_stream_b200 import MergeOpts, PsdCascade  # noqa: E402

sizes = [int(v) for v in sys.argv[1:]] or [64, 128, 256, 512, 1024, 2048, 4096, 8192]
x = (torch.rand(200_000_000, device="cuda") - 0.5) * (12 ** 0.5)
for n in sizes:
    c = PsdCascade(n)
    c.profile_enable(True)
    for _ in range(2):
        c.process(x)
    c.psd(MergeOpts())
    c.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        c.process(x)
    p, b = c.psd(MergeOpts())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    prof, launches = c.profile_read()
    print("N=%5d  %8.1f GS/s  %.3f ms/step  stages=%d  " % (n, 200e6 / ms / 1e6, ms, len(b)) +
          "  ".join("%s=%.3f" % (k, v[0] / 5) for k, v in prof.items()))
