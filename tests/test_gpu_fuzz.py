"""Seeded random walks over the public API (process host/device with ragged sizes, option changes, readouts,
clone, reset) checked against the CPU oracle after every readout: exercises the host state machine
(carry/fresh buffers, staging, second stream, EWMA plans) far from the happy path."""
import numpy as np
import pytest

from conftest import uniform_noise

pytestmark = pytest.mark.gpu


def check(g, o, sp, what):
    p, b = g.psd(sp.MergeOpts(keep_overlap=True, min_count=0, keep_transition_band=True))
    po, bo = o.psd(True, 0, True)
    got = [(k.start, k.include, k.count, k.avg, k.bins.start, k.bins.stop, k.decimation, k.pending, k.processed) for k in b]
    want = [(k.start, bool(k.include), k.count, k.avg, k.bins_start, k.bins_end, k.decimation, k.pending, k.processed) for k in bo]
    assert got == want, what
    for k in b:
        if k.count:
            sl = slice(k.start, k.start + len(k.bins))
            w = po[sl].astype(np.float64)
            floor = 1e-5 * np.median(w)
            rel = (np.abs(p[sl] - w) - floor) / np.maximum(w, 1e-300)
            # bins next to DC hold what is left of the (decimation-amplified) DC offset after detrending: a
            # difference of nearly equal numbers, set by how the f32 offset was rounded (tests/test_gpu_psd.py).
            # Under Detrend::Span the reference accumulates `offset += slope` sequentially in f32
            # (psd.rs:98-101): inside a binade every step adds the slope rounded to a multiple of ulp(offset),
            # a systematic error of up to ulp/2 per step, i.e. a ramp of up to N ulp/2 over the segment.  The
            # device evaluates x0 + i*slope with one rounding.  The ramp leaks into bin k with amplitude ~1/k,
            # so against a single noisy segment the two differ by ~N^1.5 2^-24 / k relative (1e-3 at N = 8192,
            # k = 4; 2e-5 at N = 512): the tolerance widens accordingly near DC.
            kk = np.arange(rel.size, dtype=np.float64)
            tol = 1e-4 + float(k.fft_size) ** 1.5 * 2.0 ** -24 / np.maximum(kk, 1.0)
            assert np.all(rel[4:] < tol[4:]) and np.max(rel[:4]) < 2e-2, \
                "%s: stage dec=%d rel %.3g" % (what, k.decimation, np.max(rel))


def _cases():
    base = [(1, 512), (2, 4096), (3, 64), (4, 2048), (5, 512)]
    # SSPSD_FUZZ_EXTRA=k adds k more seeded walks over all FFT sizes (for soak runs; default: none)
    import os
    extra = int(os.environ.get("SSPSD_FUZZ_EXTRA", "0"))
    sizes = [64, 128, 256, 512, 1024, 2048, 4096, 8192]
    return base + [(100 + i, sizes[i % len(sizes)]) for i in range(extra)]


@pytest.mark.parametrize("seed,n", _cases())
def test_random_api_walk(oracle, seed, n):
    import torch
    import stabilizer_stream_b200 as sp
    rng = np.random.default_rng(seed)
    x = uniform_noise(260 * n * 8, 100 + seed) + np.float32(0.05)
    xd = torch.from_numpy(x).cuda()
    g = sp.PsdCascade(n, hbf=sp.Hbf(seed & 1), host_stage=1 << 13, max_batch=int(rng.integers(20 * n, 400 * n)))
    o = oracle.Cascade(n, seed & 1)
    pos = 0
    step = 0
    while pos < x.size:
        step += 1
        r = rng.random()
        if r < 0.55:
            k = int(rng.integers(0, 30 * n)) if rng.random() < 0.8 else int(rng.integers(0, 9))
            if rng.random() < 0.5:
                g.process(x[pos:pos + k])
            else:
                g.process(xd[pos:pos + k])
            o.process(x[pos:pos + k])
            pos += k
        elif r < 0.65:
            d = int(rng.integers(0, 4))
            g.set_detrend(sp.Detrend(d))
            o.set_detrend(d)
        elif r < 0.72:
            lim = int(rng.choice([2 ** 32 - 1, 3, 17, 200, 0]))
            cnt = int(rng.choice([2 ** 32 - 1, 2 ** 32 - 2, 5000, 64]))
            g.set_avg(sp.AvgOpts(limit=lim, count=cnt))
            o.set_avg(lim, cnt)
        elif r < 0.84:
            check(g, o, sp, "seed %d step %d" % (seed, step))
        elif r < 0.90:
            g = g.clone()
            o = o.clone()
        elif r < 0.93 and pos > 0:
            g.reset()
            o = oracle.Cascade(n, seed & 1)
            # reset keeps the options (Cmd::Reset re-creates the cascades and Cmd::Send re-applies them)
            o.set_detrend(0)
            g.set_detrend(sp.Detrend(0))
            g.set_avg(sp.AvgOpts())
    check(g, o, sp, "seed %d final" % seed)


def _frame_cases():
    import os
    extra = int(os.environ.get("SSPSD_FUZZ_EXTRA", "0"))
    return list(range(6)) + [1000 + i for i in range(extra)]


@pytest.mark.parametrize("seed", _frame_cases())
def test_random_frame_streams(oracle, seed):
    """random format / batch count / stride / call sizes / host or device frames / one corrupted frame:
    traces, loss counters, error code and error position must match Frame::from_bytes + Loss::update +
    traces() looped over the same bytes (bit exact)"""
    import torch
    import stabilizer_stream_b200 as sp
    from frames_util import BATCH_BYTES, make_frames, oracle_decode_stream
    rng = np.random.default_rng(seed)
    fmt = int(rng.integers(1, 5))
    batches = int(rng.integers(1, 256 if fmt != 3 else 25))
    flen = 8 + BATCH_BYTES[fmt] * batches
    if flen > 2048 and rng.random() < 0.7:  # mostly stay within the reference's 2048-byte datagrams
        batches = (2048 - 8) // BATCH_BYTES[fmt]
        flen = 8 + BATCH_BYTES[fmt] * batches
    pad = int(rng.choice([0, 0, 8, 3, 1, 640]))
    n_frames = int(rng.integers(1, 900))
    data, flen, stride, _ = make_frames(fmt, batches, n_frames, seed=seed, drop_every=int(rng.choice([0, 7, 101])),
                                        start_seq=int(rng.choice([0, 2 ** 32 - 3 * batches, 12345])), stride=flen + pad)
    data = bytearray(data)
    bad_at = None
    if rng.random() < 0.4:
        bad_at = int(rng.integers(0, n_frames))
        which = int(rng.integers(0, 3))
        off = bad_at * stride
        if which == 0:
            data[off + int(rng.integers(0, 2))] ^= 0x40          # magic
        elif which == 1:
            data[off + 2] = int(rng.choice([0, 5, 9, 255]))      # unknown format
        else:
            data[off + 3] = (data[off + 3] + 1) & 0xFF           # batch count disagrees with the payload
    data = bytes(data)
    st, nf, lo, want = oracle_decode_stream(oracle, data, flen, stride, n_frames)
    dec = sp.FrameDecoder()
    loss = sp.Loss()
    got = None
    f0 = 0
    err = None
    while f0 < n_frames and err is None:
        k = int(rng.integers(1, n_frames + 1))
        part = data[f0 * stride:min(f0 + k, n_frames) * stride]
        cnt = min(k, n_frames - f0)
        src = torch.frombuffer(bytearray(part), dtype=torch.uint8).cuda() if rng.random() < 0.5 else part
        try:
            f, traces, ok = dec.decode(src, flen, loss, frame_stride=stride, n_frames=cnt)
        except sp.DecodeError as e:
            err = (e.status, f0 + e.frames_ok)
            traces = e.traces
        if got is None and traces:
            got = [[] for _ in traces]
        for t, (_, v) in enumerate(traces):
            got[t].append(v)
        f0 += cnt
    if st != 0:
        assert err == (st, nf), (err, st, nf)
    else:
        assert err is None
    assert (loss.received, loss.dropped, loss.seq) == (lo.received, lo.dropped, lo.seq)
    if want is not None:
        for g, w in zip(got, want):
            g, w = np.concatenate(g), np.concatenate(w)
            assert g.size == w.size and np.array_equal(g.view(np.uint32), w.view(np.uint32))


def _stage_cases():
    import os
    return list(range(6)) + [3000 + i for i in range(int(os.environ.get("SSPSD_FUZZ_EXTRA", "0")))]


@pytest.mark.parametrize("seed", _stage_cases())
def test_random_single_stage_walk(oracle, seed):
    """Psd<N> / PsdStage (psd.rs:137-288) with both windows: ragged process() calls (host and device), option
    changes; the decimated output of every call, spectrum, count, gain and buf() must match"""
    import torch
    import stabilizer_stream_b200 as sp
    rng = np.random.default_rng(seed)
    n = int(rng.choice([64, 128, 256, 512, 1024, 2048, 4096, 8192]))
    window = int(rng.integers(0, 2))
    hbf = int(rng.integers(0, 2))
    x = uniform_noise(60 * n, 500 + seed) + np.float32(0.05)
    xd = torch.from_numpy(x).cuda()
    g = sp.Psd(n, sp.Window(window), hbf=sp.Hbf(hbf))
    o = oracle.Stage(n, window, hbf)
    pos = 0
    while pos < x.size:
        r = rng.random()
        if r < 0.7:
            k = int(rng.integers(0, 6 * n)) if rng.random() < 0.8 else int(rng.integers(0, 9))
            yg = g.process(xd[pos:pos + k] if rng.random() < 0.5 else x[pos:pos + k])
            yo = o.process(x[pos:pos + k])
            pos += k
            assert yg.shape == yo.shape
            np.testing.assert_allclose(yg, yo, atol=4e-6 * 8)
        elif r < 0.8:
            d = int(rng.integers(0, 4))
            g.set_detrend(sp.Detrend(d))
            o.set_detrend(d)
        elif r < 0.9:
            a = int(rng.choice([2 ** 32 - 1, 0, 1, 5, 40]))
            g.set_avg(a)
            o.set_avg(a)
        else:
            assert g.count() == o.count() and g.gain() == o.gain()
            np.testing.assert_array_equal(g.buf(), o.buf())
            if g.count():
                w = o.spectrum().astype(np.float64)
                got = g.spectrum().astype(np.float64)
                rel = (np.abs(got - w) - 1e-5 * np.median(w)) / np.maximum(w, 1e-300)
                kk = np.arange(rel.size, dtype=np.float64)
                tol = 1e-4 + float(n) ** 1.5 * 2.0 ** -24 / np.maximum(kk, 1.0)
                assert np.all(rel[4:] < tol[4:]) and np.max(rel[:4]) < 2e-2, "n=%d rel %.3g" % (n, np.max(rel))
    assert g.count() == o.count()
    np.testing.assert_array_equal(g.buf(), o.buf())
