#!/usr/bin/env python3
"""Small driver for ncu captures of the stage-0 kernels (run under gpurun, see profiles/README.md):
processes a few 2^26-sample batches through a default N=4096 cascade.  The first psd_stage_kernel /
decim8_kernel launches of each process() call are the full-size stage-0 launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from stabilizer_stream_b200 import MergeOpts, PsdCascade  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = (torch.rand(1 << 26, device="cuda") - 0.5) * (12 ** 0.5)
c = PsdCascade(n)
for _ in range(reps):
    c.process(x)
p, b = c.psd(MergeOpts())
print("stages", [(k.decimation, k.count) for k in b], "bins", p.size)
