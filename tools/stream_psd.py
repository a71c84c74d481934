#!/usr/bin/env python3
"""Headless driver in the shape of the reference's `stream_test` binary (src/bin/stream_test.rs:35-76) on top
of the GPU cascade: read a source, feed one PsdCascade per trace, then log breaks, PSD and the FVAR sweep.

Sources (reference src/source.rs:16-48, 102-167):
  --raw FILE          native-endian f32 samples, no header (what stream_to_raw writes, stream_to_raw.rs:20-26)
  --file FILE         stabilizer frames of --frame-size bytes each (default 8 + 30*2*6*4 like source.rs:29-31)
  --noise N           synthetic power-law noise f^N generated on the device (|N| integrators/differentiators,
                      source.rs:104-118)
  --udp IP:PORT       live Stabilizer stream (source.rs:81-93, 159-165): recvmmsg batches into page-locked slots;
                      ends after --samples items per trace or 1 s without traffic
  --dsm FTW           MASH-1-1-1 modulated sine marker generated on the device (source.rs:119-130)
--repeat wraps files around; the run ends after --samples items per trace (files: at EOF without --repeat).
Host reads go through a pinned double buffer, so the H2D copy of block b+1 overlaps the kernels of block b.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from stabilizer_stream_b200 import (Break, Detrend, FrameDecoder, Loss, MergeOpts, PsdCascade, Receiver, Source,
                                    Var)  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--raw")
    ap.add_argument("--file")
    ap.add_argument("--frame-size", type=int, default=8 + 30 * 2 * 6 * 4)
    ap.add_argument("--noise", type=int)
    ap.add_argument("--dsm", type=int)
    ap.add_argument("--udp")
    ap.add_argument("--repeat", action="store_true")
    ap.add_argument("--samples", type=float, default=0, help="stop after this many items per trace (0: until EOF)")
    ap.add_argument("--fft", type=int, default=512, help="FFT size N (the reference binaries use 512)")
    ap.add_argument("--detrend", default="midpoint", choices=["none", "midpoint", "span", "mean"])
    ap.add_argument("--trace", type=int, default=0)
    ap.add_argument("--block", type=int, default=1 << 24, help="items (or frames*64) read per block")
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    det = Detrend[a.detrend.upper()]
    limit = int(a.samples)
    cas, names, loss = [], [], Loss()
    total = 0
    t0 = time.perf_counter()

    def cascades(n):
        while len(cas) < n:
            c = PsdCascade(a.fft)
            c.set_detrend(det)
            cas.append(c)

    if a.noise is not None or a.dsm is not None:
        cascades(1)
        names = ["noise" if a.dsm is None else "dsm"]
        src = Source.noise(a.noise) if a.dsm is None else Source.dsm(a.dsm)
        if limit == 0:
            ap.error("--noise/--dsm need --samples")
        while total < limit:
            n = min(a.block, limit - total)
            cas[0].process_source(src, n)  # generated and consumed on the device
            total += n
    elif a.raw:
        cascades(1)
        names = ["raw"]
        pinned = [torch.empty(a.block, dtype=torch.float32, pin_memory=True) for _ in range(2)]
        with open(a.raw, "rb") as f:
            b = 0
            while True:
                buf = pinned[b & 1]
                got = f.readinto(memoryview(buf.numpy()).cast("B"))
                n = got // 4
                if n == 0:
                    if a.repeat and total:
                        f.seek(0)
                        continue
                    break
                if limit and total + n > limit:
                    n = limit - total
                cas[0].process(buf[:n])   # returns once the H2D copy is done; kernels keep running
                total += n
                b += 1
                if limit and total >= limit:
                    break
    elif a.file:
        dec = FrameDecoder()
        fpb = max(1, a.block * 2 // a.frame_size)
        with open(a.file, "rb") as f:
            while True:
                data = f.read(fpb * a.frame_size)
                nfr = len(data) // a.frame_size
                if nfr == 0:
                    if a.repeat and total:
                        f.seek(0)
                        continue
                    break
                cascades(4)
                info = dec.process_frames(cas, data[:nfr * a.frame_size], a.frame_size, loss)
                if not names:
                    from stabilizer_stream_b200.psd import TRACE_NAMES, Format
                    names = list(TRACE_NAMES[Format(info.format)])
                total += info.samples_per_trace
                if limit and total >= limit:
                    break
    elif a.udp:
        ip, _, port = a.udp.rpartition(":")
        rx = Receiver(ip or "0.0.0.0", int(port))
        dec = FrameDecoder()
        cascades(4)
        while True:
            info = rx.pump(dec, cas, loss, max_frames=1024, timeout_ms=1000)
            if info.frames_ok == 0:
                break
            if not names:
                from stabilizer_stream_b200.psd import TRACE_NAMES, Format
                names = list(TRACE_NAMES[Format(info.format)])
            total += info.samples_per_trace
            if limit and total >= limit:
                break
    else:
        ap.error("one of --raw, --file, --udp, --noise, --dsm is required")

    c = cas[min(a.trace, len(cas) - 1)]
    y, b = c.psd(MergeOpts())
    dt = time.perf_counter() - t0
    f = Break.frequencies(b)
    fdev = []
    if b:
        var = Var(dc_cut=1, clip=1.0)          # stream_test.rs:62
        tau = 1.0
        while tau <= b[0].effective_fft_size() // 2:
            fdev.append((tau, float(np.sqrt(max(var.eval(y, f, tau), 0.0)))))
            tau *= 2.0
    out = {"trace": names[min(a.trace, len(names) - 1)] if names else None, "items_per_trace": int(total),
           "traces": len(cas), "seconds": dt, "MSps_per_trace": total / dt / 1e6,
           "breaks": [{"start": k.start, "count": k.count, "bins": [k.bins.start, k.bins.stop], "decimation": k.decimation,
                       "pending": k.pending, "processed": k.processed, "include": k.include} for k in b],
           "psd_bins": int(y.size), "psd_median": float(np.median(y)) if y.size else None,
           "fdev": fdev[:8], "loss": {"received": loss.received, "dropped": loss.dropped}}
    if a.json:
        print(json.dumps(out))
    else:
        print("breaks:", out["breaks"])
        print("psd: %d bins, median %.6g" % (out["psd_bins"], out["psd_median"] or 0))
        print("fdev:", out["fdev"])
        if loss.received:
            print("Loss: %g %% (%d of %d)" % (100 * loss.ratio(), loss.dropped, loss.received + loss.dropped))
        print("%.1f MS/s per trace over %.2f s" % (out["MSps_per_trace"], dt))


if __name__ == "__main__":
    main()
