"""Independent float64 numpy/scipy model of the cascaded PSD (TEST INFRASTRUCTURE ONLY).

Written from the prose semantics in SURVEY.md App. A (reference src/psd.rs:196-269, 456-468,
479-543), deliberately NOT sharing code with sspsd_oracle.c: whole-stream vectorised maths
(strided segments + rfft, lfilter for the half-band FIRs) instead of the streaming state machine.
It is the accuracy "truth" for tolerance statements and cross-validates the C oracle.
"""
import os
import re

import numpy as np
from scipy import signal

_HERE = os.path.dirname(os.path.abspath(__file__))
DEPTH = 3


def load_taps():
    """Parse oracle/hbf_taps.h -> (taps[preset][set] as float64 arrays, drain[preset])."""
    src = open(os.path.join(_HERE, "hbf_taps.h")).read()
    ntaps = [[int(v) for v in grp.split(",")] for grp in
             re.search(r"orc_hbf_ntaps\[2\]\[3\] = \{\{(.*?)\}, \{(.*?)\}\};", src).groups()]
    drain = [int(v) for v in re.search(r"orc_hbf_drain\[2\] = \{(.*?)\};", src).group(1).split(",")]
    body = src[src.index("orc_hbf_taps[2][3]"):]
    rows = re.findall(r"\{([^{}]+)\}", body)
    taps = []
    for p in range(2):
        sets = []
        for s in range(3):
            vals = [float(v.strip().rstrip("f")) for v in rows[3 * p + s].split(",")]
            sets.append(np.array(vals[:ntaps[p][s]], dtype=np.float32).astype(np.float64))
        taps.append(sets)
    return taps, drain


TAPS, DRAIN = load_taps()


def hbf_impulse(t):
    """Full causal half-band impulse response (length 4M-1) from the M unique taps."""
    m = len(t)
    h = np.zeros(4 * m - 1)
    c = 2 * m - 1
    h[c] = 1.0  # taps are 2*remez(...): DC gain 2 per half-band stage (see sspsd_oracle.c)
    for i, tk in enumerate(t):
        k = m - 1 - i
        h[c - (2 * k + 1)] = tk
        h[c + (2 * k + 1)] = tk
    return h


def hbf8(x, preset):
    """Zero-state divide-by-8: three half-band FIRs, highest rate first (tap sets 2, 1, 0)."""
    y = np.asarray(x, dtype=np.float64)
    for s in (2, 1, 0):
        h = hbf_impulse(TAPS[preset][s])
        full = signal.lfilter(h, [1.0], y)
        y = full[1::2]
    return y


def window(n, kind):
    if kind == 0:
        return np.ones(n), 1.0, 1.0, 0
    return np.sin(np.pi * np.arange(n) / n) ** 2, 0.25, 1.5, n // 2


def detrend(seg, mode):
    n = seg.shape[-1]
    if mode == 0:
        return seg
    if mode == 1:
        return seg - seg[..., n // 2:n // 2 + 1]
    if mode == 2:
        slope = (seg[..., -1:] - seg[..., :1]) / (n - 1)
        return seg - (seg[..., :1] + slope * np.arange(n))
    if mode == 3:
        return seg - seg.mean(axis=-1, keepdims=True)
    raise NotImplementedError


def stage(x, n, win_kind=1, det=0, avg=2 ** 32 - 1, preset=1, count0=0, spec0=None):
    """One Psd<N> stage over the whole stream x (zero initial state).
    Returns dict(spectrum, count, count_raw, decimated (stream for the next stage), pending)."""
    x = np.asarray(x, dtype=np.float64)
    w, power, nenbw, overlap = window(n, win_kind)
    hop = n - overlap
    L = x.size
    craw = 0 if L < n else 1 + (L - n) // hop
    spec = np.zeros(n // 2 + 1) if spec0 is None else spec0.copy()
    count = count0
    if craw:
        idx = np.arange(n)[None, :] + hop * np.arange(craw)[:, None]
        X = np.fft.rfft(detrend(x[idx], det) * w, axis=-1)
        P = X.real ** 2 + X.imag ** 2
        if avg >= count + craw:
            spec += P.sum(axis=0)
            count += craw
        else:
            for k in range(craw):
                if count > avg:
                    g = np.float64(np.float32(avg) / np.float32(count))
                    count = avg
                else:
                    g = 1.0
                count += 1
                spec = g * spec + P[k]
    D = n + (craw - 1) * hop if craw else 0
    dec = hbf8(x[:D], preset)[DRAIN[preset]:] if D else np.zeros(0)
    pending = L - craw * hop if craw else L
    return dict(spectrum=spec, count=count, count_raw=craw, decimated=dec, pending=pending,
                gain=(n // 2) * count * nenbw * power, overlap=overlap)


def cascade(x, n, det=0, avg_limit=2 ** 32 - 1, avg_count=2 ** 32 - 1, preset=1):
    """PsdCascade<N> over the whole stream: list of per-stage dicts (stage 0 first)."""
    out = []
    s = np.asarray(x, dtype=np.float64)
    i = 0
    while s.size > 0:
        avg = min(avg_count >> (DEPTH * i) if DEPTH * i < 32 else 0, avg_limit)
        st = stage(s, n, 1, det, avg, preset)
        st["avg"] = avg
        out.append(st)
        s = st["decimated"]
        i += 1
    return out


def merge(stages, n, keep_overlap=False, min_count=1, keep_transition_band=False):
    """PsdCascade::psd: returns (p, breaks as dicts)."""
    p = []
    breaks = []
    decimation = 1 << (DEPTH * len(stages))
    end = 0
    for st in reversed(stages):
        decimation >>= DEPTH
        start = 0 if keep_overlap else (end + 7) >> 3
        end = 2 * n // 5 if (decimation > 1 and not keep_transition_band) else n // 2 + 1
        include = st["count"] >= min_count
        cnt = st["count"]
        breaks.append(dict(start=len(p), include=include, count=cnt, avg=st["avg"], bins=(start, end),
                           fft_size=n, decimation=decimation, pending=st["pending"],
                           processed=n * cnt - st["overlap"] * max(cnt - 1, 0)))
        if include:
            g = 1.0 / (st["gain"] * decimation)
            p.extend(st["spectrum"][start:end] * g)
        else:
            end = start
    return np.array(p), breaks
