#!/usr/bin/env python3
"""A/B timing of kernel / scheduling variants selected by environment switches that the library reads
when a handle is created (SSPSD_K2, SSPSD_K3, SSPSD_DEFER, SSPSD_OVERLAP, ...).  Device-resident
200e6-sample steps of the default N=4096 cascade, CUDA events, one JSON line per variant.

    python tools/bench_variants.py "SSPSD_K3=tiled" "SSPSD_K3=tma960" "SSPSD_K3=tma640,SSPSD_DEFER=1" ...
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from stabilizer_stream_b200 import MergeOpts, PsdCascade  # noqa: E402

N = int(os.environ.get("BENCH_N", "4096"))
SAMPLES = 200_000_000
STEPS = int(os.environ.get("BENCH_STEPS", "10"))
x = (torch.rand(SAMPLES, device="cuda") - 0.5) * (12 ** 0.5)
stream = torch.cuda.current_stream()


def run(readout_every_step, profile):
    c = PsdCascade(N, stream=stream.cuda_stream or 1)
    c.profile_enable(profile)
    for _ in range(3):
        c.process(x)
        c.psd(MergeOpts())
    c.profile_read()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(STEPS):
        c.process(x)
        if readout_every_step:
            c.psd(MergeOpts())
    p, b = c.psd(MergeOpts())
    e1.record(stream)
    torch.cuda.synchronize()
    prof, launches = c.profile_read()
    return e0.elapsed_time(e1) / STEPS, prof, launches, float(p.sum())


for spec in sys.argv[1:] or [""]:
    keys = []
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
        keys.append(k)
    ms, _, launches, chk = run(False, False)
    ms_rs, _, launches_rs, _ = run(True, False)
    _, prof, _, _ = run(False, True)
    out = {"variant": spec or "default", "n_fft": N, "ms_per_step": ms, "GSps": SAMPLES / ms / 1e6,
           "ms_per_step_readout_every_step": ms_rs, "GSps_readout_every_step": SAMPLES / ms_rs / 1e6,
           "launches_per_step": launches / STEPS, "launches_per_step_readout": launches_rs / STEPS, "checksum": chk,
           "kernel_ms_per_step": {k: round(v[0] / STEPS, 4) for k, v in prof.items()},
           "kernel_launches": {k: v[1] for k, v in prof.items()}}
    print(json.dumps(out), flush=True)
    for k in keys:
        del os.environ[k]
