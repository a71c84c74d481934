#!/usr/bin/env python3
"""Throughput of the device-resident synthetic sources (source.rs:104-134 on the GPU) and of a cascade
fed by them without touching host memory."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import stabilizer_stream_b200 as m  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    n = 200_000_000
    buf = torch.empty(n, dtype=torch.float32, device="cuda")
    out = {}
    for name, make in [("noise 0", lambda: m.Source.noise(0)), ("noise +1", lambda: m.Source.noise(1)),
                       ("noise +2", lambda: m.Source.noise(2)), ("noise -1", lambda: m.Source.noise(-1)),
                       ("noise -2", lambda: m.Source.noise(-2)), ("noise -4", lambda: m.Source.noise(-4)),
                       ("dsm", lambda: m.Source.dsm(0x1234567))]:
        src = make()
        ms = timed(lambda: src.get(n, out=buf))
        out[name] = {"ms": round(ms, 3), "GSps": round(n / ms / 1e6, 1), "GBps_written": round(4 * n / ms / 1e6, 1)}
    for name, make in [("noise 0", lambda: m.Source.noise(0)), ("noise -1", lambda: m.Source.noise(-1))]:
        c = m.PsdCascade(4096)
        src = make()
        ms = timed(lambda: c.process_source(src, n))
        out["cascade N=4096 <- " + name] = {"ms": round(ms, 3), "GSps": round(n / ms / 1e6, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
