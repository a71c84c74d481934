"""Bit-exact parity of the CUDA frame decoder + loss accounting (through the C ABI) with the golden
fixtures and with the CPU oracle on synthetic frame streams.  Needs a B200."""
import json
import os

import numpy as np
import pytest

from frames_util import make_frames, oracle_decode_stream

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(G, "frames.json")))


@pytest.fixture(scope="module")
def sp():
    import torch
    assert torch.cuda.is_available()
    import stabilizer_stream_b200 as m
    return m


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_frames(sp, name):
    c = CASES[name]
    data = open(os.path.join(G, "frames_%s.bin" % name), "rb").read()
    loss = sp.Loss()
    fmt, traces, ok = sp.FrameDecoder().decode(data, c["frame_len"], loss)
    assert ok == 1 and int(fmt) == c["format"]
    assert [n for n, _ in traces] == c["names"]
    for (_, got), want in zip(traces, c["expect_bits"]):
        assert got.view(np.uint32).tolist() == want
    assert (loss.received, loss.dropped, loss.seq) == (c["batches"], 0, (c["seq"] + c["batches"]) & 0xFFFFFFFF)


@pytest.mark.parametrize("fmt,batches,stride_pad", [(1, 22, 0), (1, 22, 8), (1, 7, 3), (2, 25, 0), (3, 18, 0), (4, 60, 0),
                                                    (1, 255, 0), (2, 1, 4)])
@pytest.mark.parametrize("device_frames", [False, True])
def test_streams_match_oracle_bit_exact(sp, oracle, fmt, batches, stride_pad, device_frames):
    import torch
    n_frames = 700
    flen = 8 + {1: 64, 2: 56, 3: 80, 4: 24}[fmt] * batches
    data, flen, stride, hdrs = make_frames(fmt, batches, n_frames, seed=fmt * 100 + batches, drop_every=13,
                                           start_seq=0xFFFFFFFF - 40 * batches, stride=flen + stride_pad)
    st, nf, lo, want = oracle_decode_stream(oracle, data, flen, stride, n_frames)
    assert st == 0
    dec = sp.FrameDecoder()
    loss = sp.Loss()
    half = (n_frames // 2) * stride
    parts = None
    for part in (data[:half], data[half:]):   # loss state carries across calls
        src = torch.frombuffer(bytearray(part), dtype=torch.uint8).cuda() if device_frames else part
        f, traces, ok = dec.decode(src, flen, loss, frame_stride=stride)
        assert int(f) == fmt
        if parts is None:
            parts = [[] for _ in traces]
        for t, (_, v) in enumerate(traces):
            parts[t].append(v)
    assert (loss.received, loss.dropped, loss.seq) == (lo.received, lo.dropped, lo.seq)
    assert lo.dropped > 0
    assert len(parts) == len(want)
    for g, w in zip(parts, want):
        g = np.concatenate(g)
        w = np.concatenate(w)
        assert g.size == w.size and np.array_equal(g.view(np.uint32), w.view(np.uint32))


def test_malformed_frames(sp, oracle):
    from stabilizer_stream_b200 import _lib
    data, flen, stride, hdrs = make_frames(1, 4, 20, seed=1)
    dec = sp.FrameDecoder()
    for pos, patch, code in ((5, (0, b"\x00"), _lib.EHEADER), (9, (2, b"\x07"), _lib.EFORMAT), (0, (3, b"\x05"), _lib.EBATCHES),
                             (19, (2, b"\x02"), _lib.EFORMAT)):
        bad = bytearray(data)
        bad[pos * stride + patch[0]:pos * stride + patch[0] + 1] = patch[1]
        st, nf, lo, want = oracle_decode_stream(oracle, bytes(bad), flen, stride, 20)
        loss = sp.Loss()
        with pytest.raises(sp.DecodeError) as e:
            dec.decode(bytes(bad), flen, loss)
        if pos == 19 and patch[0] == 2:
            # a *valid* other format mid-batch is rejected by the batched API (one format per call)
            assert e.value.status == _lib.EFORMAT and e.value.frames_ok == 19
            continue
        assert e.value.status == st and e.value.frames_ok == nf == pos
        assert (loss.received, loss.dropped, loss.seq) == (lo.received, lo.dropped, lo.seq)
        if pos:
            for (_, g), w in zip(e.value.traces, want):
                assert np.array_equal(g.view(np.uint32), np.concatenate(w).view(np.uint32))
    # payload size not a multiple of the batch size
    with pytest.raises(sp.DecodeError) as e:
        dec.decode(data[:stride - 1], flen - 1, sp.Loss())
    assert e.value.status == _lib.ESIZE
    with pytest.raises(sp.DecodeError) as e:
        dec.decode(data[:7], 7, sp.Loss(), n_frames=1)
    assert e.value.status == _lib.ESHORT
    # empty input
    f, traces, ok = dec.decode(b"", flen, sp.Loss())
    assert ok == 0 and traces == []


def test_frames_into_cascades(sp, oracle):
    """config 3: AdcDac frames decoded on the device straight into four cascades + loss."""
    n = 512
    data, flen, stride, hdrs = make_frames(1, 22, 3000, seed=9, drop_every=1009, start_seq=0xFFFFFF00)
    st, nf, lo, want = oracle_decode_stream(oracle, data, flen, stride, 3000)
    cas = [sp.PsdCascade(n) for _ in range(4)]
    ocs = [oracle.Cascade(n, 1) for _ in range(4)]
    for c, o in zip(cas, ocs):
        c.set_detrend(sp.Detrend.MIDPOINT)   # stream_test.rs:41
        o.set_detrend(1)
    dec = sp.FrameDecoder()
    loss = sp.Loss()
    third = 1000 * stride
    for k in range(3):
        info = dec.process_frames(cas, data[k * third:(k + 1) * third], flen, loss)
        assert info.frames_ok == 1000 and info.n_traces == 4 and info.samples_per_trace == 1000 * 22 * 8
    assert (loss.received, loss.dropped, loss.seq) == (lo.received, lo.dropped, lo.seq)
    for t in range(4):
        ocs[t].process(np.concatenate(want[t]))
        p, b = cas[t].psd()
        po, bo = ocs[t].psd()
        assert [k.count for k in b] == [k.count for k in bo]
        assert p.size == po.size and np.max(np.abs(p - po) / np.maximum(po, 1e-30)) < 1e-4


@pytest.mark.parametrize("fmt,batches", [(2, 25), (3, 18), (4, 60)])
def test_low_rate_formats_into_cascades(sp, oracle, fmt, batches):
    """Fls / ThermostatEem / Mpll frames (one item per batch per trace) decoded on the device into cascades."""
    n = 64
    data, flen, stride, hdrs = make_frames(fmt, batches, 1500, seed=20 + fmt, drop_every=211)
    st, nf, lo, want = oracle_decode_stream(oracle, data, flen, stride, 1500)
    ntr = len(want)
    cas = [sp.PsdCascade(n) for _ in range(ntr)]
    loss = sp.Loss()
    dec = sp.FrameDecoder()
    info = dec.process_frames(cas, data, flen, loss)
    assert info.n_traces == ntr and info.samples_per_trace == 1500 * batches
    assert (loss.received, loss.dropped, loss.seq) == (lo.received, lo.dropped, lo.seq)
    for t in range(ntr):
        w = np.concatenate(want[t])
        if not np.all(np.isfinite(w)) or np.max(np.abs(w)) > 1e30:
            continue   # random f32 payload words (ThermostatEem) may be NaN/huge: decode parity is covered bit exactly above
        o = oracle.Cascade(n, 1)
        o.process(w)
        p, b = cas[t].psd()
        po, bo = o.psd()
        assert [k.count for k in b] == [k.count for k in bo]
        scale = np.median(po)
        assert np.max((np.abs(p - po) - 1e-5 * scale) / np.maximum(po, 1e-300)) < 1e-4


def test_eshort_retry_accounts_the_batch_once(sp, oracle):
    """Capacity-in / length-out retry (include/sspsd.h): a too small host trace buffer returns ESHORT with the
    needed length and must leave `Loss` untouched, so that the retried call counts the batch exactly once
    (the first version advanced received/dropped/seq before the capacity check)."""
    import ctypes as C

    from stabilizer_stream_b200 import _lib as L
    n_frames, batches = 50, 22
    data, flen, stride, hdrs = make_frames(1, batches, n_frames, seed=77, drop_every=7, start_seq=0xFFFFFF00)
    st, nf, lo, want = oracle_decode_stream(oracle, data, flen, stride, n_frames)
    dec = sp.FrameDecoder()
    loss = sp.Loss()
    loss.update(0xFFFFFF00 - 3 * batches, batches)   # some history, so `dropped` moves on the first frame
    before = (loss.received, loss.dropped, loss.seq)
    buf = np.frombuffer(data, np.uint8)
    need = n_frames * batches * 8
    small = [np.zeros(need - 1, np.float32) for _ in range(4)]
    ptrs = (C.c_void_p * 4)(*[t.ctypes.data for t in small])
    info = L.DecodeInfoC()
    rc = L.lib().sspsd_decode_frames(dec._h, buf.ctypes.data, n_frames, flen, stride, L.MEM_HOST, C.byref(loss.c), ptrs,
                                     need - 1, L.MEM_HOST, C.byref(info))
    assert rc == L.ESHORT and info.samples_per_trace == need and info.frames_ok == 0
    assert (loss.received, loss.dropped, loss.seq) == before
    big = [np.zeros(need, np.float32) for _ in range(4)]
    ptrs = (C.c_void_p * 4)(*[t.ctypes.data for t in big])
    rc = L.lib().sspsd_decode_frames(dec._h, buf.ctypes.data, n_frames, flen, stride, L.MEM_HOST, C.byref(loss.c), ptrs,
                                     need, L.MEM_HOST, C.byref(info))
    assert rc == L.OK and info.frames_ok == n_frames
    ref = oracle.Loss()
    ref.update(0xFFFFFF00 - 3 * batches, batches)
    for seq, b in hdrs:
        ref.update(seq, b)
    assert (loss.received, loss.dropped, loss.seq) == (ref.received, ref.dropped, ref.seq)
    for g, w in zip(big, want):
        assert np.array_equal(g.view(np.uint32), np.concatenate(w).view(np.uint32))


@pytest.mark.parametrize("fmt,batches,bad_at", [(1, 3, None), (1, 3, 30_000), (2, 1, None), (2, 1, 16_667), (4, 2, 49_999)])
def test_large_host_batches_are_pipelined_in_sub_chunks(sp, oracle, fmt, batches, bad_at):
    """>= 32768 host frames per sspsd_cascade_process_frames call are cut into sub-chunks (copy k + 1 overlaps decode and
    cascades of k); the loss state chains from sub-chunk to sub-chunk on the host, the device's counters are the ones
    returned.  Against the frame-by-frame oracle, with drops, a sequence wrap, trace offsets that are not 16-byte aligned
    (fmt 2, one item per frame) and a malformed frame in the middle / at the start of a sub-chunk / in the last frame."""
    n, n_frames = 512, 50_000
    data, flen, stride, hdrs = make_frames(fmt, batches, n_frames, seed=77 + fmt, drop_every=997,
                                           start_seq=0xFFFFFFFF - 20_000 * batches)
    if bad_at is not None:
        b = bytearray(data)
        b[bad_at * stride] = 0x7A   # magic, frame.rs:26-28
        data = bytes(b)
    st, nf, lo, want = oracle_decode_stream(oracle, data, flen, stride, n_frames)
    assert nf == (n_frames if bad_at is None else bad_at)
    ntr = 3 if fmt == 4 else 4
    cas = [sp.PsdCascade(n) for _ in range(ntr)]
    dec = sp.FrameDecoder()
    loss = sp.Loss()
    if bad_at is None:
        info = dec.process_frames(cas, data, flen, loss)
        assert info.frames_ok == n_frames
    else:
        with pytest.raises(sp.DecodeError) as e:
            dec.process_frames(cas, data, flen, loss)
        assert e.value.frames_ok == bad_at
    assert (loss.received, loss.dropped, loss.seq) == (lo.received, lo.dropped, lo.seq)
    assert lo.dropped > 0
    # a second call continues the stream (state carried by `loss` and the cascades)
    tail, _, _, _ = make_frames(fmt, batches, 100, seed=5, start_seq=(lo.seq + 7 * batches) & 0xFFFFFFFF)
    dec.process_frames(cas, tail, flen, loss)
    st2, nf2, lo2, want2 = oracle_decode_stream(oracle, tail, flen, stride, 100)
    assert loss.received == lo.received + lo2.received and loss.dropped == lo.dropped + lo2.dropped + 7 * batches
    for t in range(ntr):
        o = oracle.Cascade(n, 1)
        o.process(np.concatenate(want[t] + want2[t]))
        p, b = cas[t].psd()
        po, bo = o.psd()
        assert [k.count for k in b] == [k.count for k in bo]
        assert p.size == po.size and np.max(np.abs(p - po) / np.maximum(po, 1e-30)) < 1e-4
