"""UDP ingest -> pinned slot ring -> batched decode -> cascades (source.rs:159-165 + bin/psd.rs:174-182 in
one call) against the same frames decoded from memory.  Needs a B200."""
import socket

import numpy as np
import pytest

from frames_util import make_frames, oracle_decode_stream

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import torch
    assert torch.cuda.is_available()
    import stabilizer_stream_b200 as m
    return m


def test_udp_pump_matches_decoding_from_memory(sp, oracle):
    n_frames, batches = 600, 22
    data, flen, stride, hdrs = make_frames(1, batches, n_frames, seed=11, drop_every=97, start_seq=2 ** 32 - 5000)
    rx = sp.Receiver("127.0.0.1", 0, n_slots=256)  # page-locked slots
    tx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    tx.connect(("127.0.0.1", rx.port))
    dec = sp.FrameDecoder()
    cas = [sp.PsdCascade(512) for _ in range(4)]
    loss = sp.Loss()
    done = 0
    for base in range(0, n_frames, 16):
        for f in range(base, min(base + 16, n_frames)):
            tx.send(data[f * stride:f * stride + flen])
        want = min(base + 16, n_frames)
        while done < want:
            info = rx.pump(dec, cas, loss, max_frames=256, timeout_ms=2000)
            assert info.frames_ok > 0 and info.format == 1 and info.n_traces == 4
            done += info.frames_ok
    assert rx.datagrams == n_frames

    # the same stream decoded from memory in one batch
    ref = [sp.PsdCascade(512) for _ in range(4)]
    rloss = sp.Loss()
    sp.FrameDecoder().process_frames(ref, data, flen, rloss)
    assert (loss.received, loss.dropped, loss.seq) == (rloss.received, rloss.dropped, rloss.seq)
    # and the loss counters of the CPU restatement, bit for bit
    st, ok, oloss, _ = oracle_decode_stream(oracle, data, flen, stride, n_frames)
    assert st == 0 and (loss.received, loss.dropped) == (oloss.c.received, oloss.c.dropped)
    for a, b in zip(cas, ref):
        pa, ba = a.psd()
        pb, bb = b.psd()
        assert [(k.count, k.pending) for k in ba] == [(k.count, k.pending) for k in bb]
        np.testing.assert_allclose(pa, pb, rtol=1e-5)


def test_malformed_datagram_raises_like_from_bytes(sp):
    rx = sp.Receiver("127.0.0.1", 0, n_slots=16)
    tx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    tx.connect(("127.0.0.1", rx.port))
    tx.send(bytes([0x7B, 0x06, 1, 1]) + bytes(4 + 64))  # wrong magic: de::Error::InvalidHeader
    with pytest.raises(sp.DecodeError) as e:
        rx.pump(sp.FrameDecoder(), [sp.PsdCascade(512)], sp.Loss(), timeout_ms=2000)
    assert e.value.status == 5
