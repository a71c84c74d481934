#!/usr/bin/env python3
"""Driver for ncu captures of the frame-decode kernels: 2^18 AdcDac frames (22 batches) decoded to device traces."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
from frames_util import make_frames  # noqa: E402

from stabilizer_stream_b200 import FrameDecoder, Loss  # noqa: E402

small, flen, stride, _ = make_frames(1, 22, 4096, seed=3, drop_every=1009)
data = small * 64
fr = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
nfr = len(data) // stride
outs = [torch.empty(nfr * 22 * 8, device="cuda") for _ in range(4)]
dec = FrameDecoder()
for _ in range(3):
    info = dec.decode_device(fr, flen, outs, Loss())
torch.cuda.synchronize()
print("frames", nfr, "bytes", len(data), "samples/trace", info.samples_per_trace)
