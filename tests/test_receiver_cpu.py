"""UDP receiver (source.rs:81-93, 159-165 replaced by recvmmsg into a slot ring) over loopback.
Host logic only: the slots live in plain host memory here (no CUDA device needed)."""
import socket

import numpy as np
import pytest

from frames_util import make_frames


@pytest.fixture()
def rx():
    import stabilizer_stream_b200 as m
    r = m.Receiver("127.0.0.1", 0, pinned=False, n_slots=64)
    s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    s.connect(("127.0.0.1", r.port))
    yield r, s
    s.close()


def test_timeout_returns_no_frames(rx):
    r, _ = rx
    frames, ln = r.recv(16, timeout_ms=50)
    assert frames.shape == (0, 2048) and ln == 0
    assert r.datagrams == 0


def test_datagrams_arrive_in_order_and_intact(rx):
    r, s = rx
    data, flen, stride, hdrs = make_frames(1, 22, 40, seed=3)
    sent = [data[i * stride:i * stride + flen] for i in range(40)]
    got = []
    for base in range(0, 40, 8):  # small bursts: the default socket buffer holds only ~100 KB
        for f in sent[base:base + 8]:
            s.send(f)
        while len(got) < base + 8:
            frames, ln = r.recv(64, timeout_ms=2000)
            assert ln == flen and frames.shape[1] == r.slot_bytes == 2048
            got += [bytes(row[:ln]) for row in frames]
    assert got == sent
    assert r.datagrams == 40


def test_runs_split_where_the_datagram_size_changes(rx):
    r, s = rx
    a = [bytes([i]) * 100 for i in range(3)]
    b = [bytes([9]) * 64]
    c = [bytes([7]) * 100]
    for f in a + b + c:
        s.send(f)
    runs = []
    while sum(n for n, _ in runs) < 5:
        frames, ln = r.recv(64, timeout_ms=2000)
        assert frames.shape[0] > 0
        runs.append((frames.shape[0], ln))
        for row in frames:
            assert len(set(row[:ln].tolist())) == 1
    # every run is homogeneous, and the size sequence is preserved
    assert [ln for n, ln in runs for _ in range(n)] == [100, 100, 100, 64, 100]


def test_max_frames_bounds_a_run_and_keeps_the_rest(rx):
    r, s = rx
    for i in range(6):
        s.send(bytes([i]) * 32)
    first = []
    total = 0
    while total < 6:
        frames, ln = r.recv(4, timeout_ms=2000)
        assert 0 < frames.shape[0] <= 4 and ln == 32
        first += [int(row[0]) for row in frames]
        total += frames.shape[0]
    assert first == list(range(6))


def test_oversized_datagram_is_cut_to_the_slot(rx):
    r, s = rx
    s.send(bytes(3000))
    frames, ln = r.recv(4, timeout_ms=2000)
    assert frames.shape[0] == 1 and ln == 2048  # like the reference's 2048-byte read buffer, source.rs:160


def test_bad_arguments():
    import stabilizer_stream_b200 as m
    with pytest.raises(m.psd.L.SspsdError):
        m.Receiver("not-an-ip", 0, pinned=False)
    with pytest.raises(m.psd.L.SspsdError):
        m.Receiver("127.0.0.1", 0, slot_bytes=1001, pinned=False)
