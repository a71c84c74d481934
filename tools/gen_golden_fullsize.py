#!/usr/bin/env python3
"""Golden per-stage spectra of the BASELINE configs at FULL size.

Runs, in this (CPU) container, the f32 restatement of the reference (oracle/sspsd_oracle.c) and the float64
truth model (oracle/model_f64.c) over the counter-based synthetic stream the device generates bit-exactly
(Philox4x32-10 white noise, seed 0x7654321: SourceOpts --noise 0, reference src/source.rs:104-117), and
stores per stage: the f32 oracle's accumulator row, its averaging count, the f64 row, plus the oracle's
merged psd() and breaks.  tests/test_gpu_fullsize.py regenerates the same stream on the GPU
(sspsd_cascade_process_source) and compares against these files, so nothing large crosses PCIe or the
repo.  The f32 oracle's own drift against f64 (reference src/psd.rs:171-172) is stored too.

    python tools/gen_golden_fullsize.py [name ...]      # default: all configs; minutes of CPU time
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as orc  # noqa: E402

SEED = 0x7654321
BLOCK = 1 << 24
U32MAX = 0xFFFFFFFF
# name -> (n_fft, total samples, single stage?, detrend, (avg_limit, avg_count))
CONFIGS = {
    "c1_single_none": (4096, 1 << 28, True, orc.DETREND_NONE, (U32MAX, U32MAX)),
    "c1_single_mean": (4096, 1 << 28, True, orc.DETREND_MEAN, (U32MAX, U32MAX)),
    "c2_default": (4096, 200_000_000, False, orc.DETREND_NONE, (U32MAX, U32MAX)),
    # the psd binary's preset: Detrend::Mean, AvgOpts{limit: avg_max - 1, count: avg - 1}, src/bin/psd.rs:41-78
    "c2_preset": (4096, 200_000_000, False, orc.DETREND_MEAN, (999, U32MAX - 1)),
    "c2_n512": (512, 200_000_000, False, orc.DETREND_NONE, (U32MAX, U32MAX)),
    "c5_default": (4096, 4_800_000_000, False, orc.DETREND_NONE, (U32MAX, U32MAX)),
    "c5_preset": (4096, 4_800_000_000, False, orc.DETREND_MEAN, (999, U32MAX - 1)),
}


def run(name):
    n, total, single, det, (lim, cnt) = CONFIGS[name]
    src = orc.Source(orc.SOURCE_NOISE, 0, SEED)
    if single:
        o32 = orc.Stage(n, orc.WINDOW_HANN, orc.HBF_140)
        o32.set_detrend(det)
    else:
        o32 = orc.Cascade(n, orc.HBF_140)
        o32.set_detrend(det)
        o32.set_avg(lim, cnt)
    o64 = orc.CascadeF64(n, orc.HBF_140, orc.WINDOW_HANN, det, lim, cnt, 1 if single else 0)
    t32 = t64 = tg = 0.0
    pos = 0
    while pos < total:
        m = min(BLOCK, total - pos)
        t0 = time.perf_counter()
        x = src.get(m)
        t1 = time.perf_counter()
        o32.process(x)
        t2 = time.perf_counter()
        o64.process(x)
        t3 = time.perf_counter()
        tg += t1 - t0
        t32 += t2 - t1
        t64 += t3 - t2
        pos += m
    out = {}
    if single:
        rows32, counts = [o32.spectrum()], [o32.count()]
        gains = [o32.gain()]
    else:
        ns = o32.num_stages()
        rows32 = [o32.stage_spectrum(i) for i in range(ns)]
        counts = [o32.stage_count(i) for i in range(ns)]
        p, b = o32.psd()
        out["psd"] = p
        out["breaks"] = np.array([k.as_tuple() for k in b], dtype=np.uint64)
        gains = []
    ns64 = o64.num_stages()
    assert ns64 == len(rows32), (ns64, len(rows32))
    rows64, c64 = [], []
    for i in range(ns64):
        sp, c, craw, L = o64.stage(i)
        rows64.append(sp)
        c64.append(c)
    assert c64 == [int(c) for c in counts], (c64, counts)
    out["rows32"] = np.stack(rows32).astype(np.float32)
    out["rows64"] = np.stack(rows64).astype(np.float64)
    out["counts"] = np.array(counts, dtype=np.uint64)
    # the f32 restatement's own drift against f64, per stage (max over bins, relative to the f64 value)
    drift = []
    for r32, r64, c in zip(rows32, rows64, counts):
        if c == 0:
            drift.append(0.0)
            continue
        drift.append(float(np.max(np.abs(r32.astype(np.float64) - r64) / np.maximum(r64, 1e-300))))
    meta = dict(name=name, n_fft=n, total=total, single_stage=single, detrend=int(det), avg_limit=lim, avg_count=cnt,
                seed=SEED, hbf="140", counts=[int(c) for c in counts], oracle_f32_drift_vs_f64=drift,
                seconds=dict(source=tg, oracle_f32=t32, model_f64=t64),
                oracle_ms_per_s=total / t32 / 1e6)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "fullsize_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(json.dumps(meta), flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CONFIGS)):
        run(nm)
