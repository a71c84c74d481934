"""The CPU oracle must be re-entrant: bench.py --impl reference runs one cascade per thread (one per
channel, reference src/bin/psd.rs:174-182 spread over the host cores).  Round 1 kept the FFT work planes
in the process-global plan, so concurrent cascades corrupted each other."""
import threading

import numpy as np

from oracle import binding as orc


def _stream(seed, n):
    rng = np.random.default_rng(seed)
    return ((rng.random(n, dtype=np.float32) - np.float32(0.5)) * np.float32(12 ** 0.5)).astype(np.float32)


def _run(n_fft, x, out, i):
    c = orc.Cascade(n_fft, orc.HBF_140)
    for k in range(0, x.size, 1 << 16):
        c.process(x[k:k + (1 << 16)])
    out[i] = c.psd()[0]


def test_four_threads_bit_equal_to_single_thread():
    orc.lib()
    n = 1 << 21
    xs = [_stream(100 + i, n) for i in range(4)]
    sizes = [4096, 4096, 512, 1024]  # same and different FFT sizes share the plan table
    single = [None] * 4
    for i in range(4):
        _run(sizes[i], xs[i], single, i)
    multi = [None] * 4
    ths = [threading.Thread(target=_run, args=(sizes[i], xs[i], multi, i)) for i in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for i in range(4):
        assert single[i].size == multi[i].size
        assert np.array_equal(single[i].view(np.uint32), multi[i].view(np.uint32)), "channel %d differs" % i


def test_plan_creation_race():
    """First use of an FFT size from several threads at once (plan tables built under a lock)."""
    orc.lib()
    x = _stream(7, 1 << 15)
    want = orc.fft_forward(x[:2048].astype(np.complex64))
    got = [None] * 8

    def work(i):
        got[i] = orc.fft_forward(x[:2048].astype(np.complex64))

    ths = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for g in got:
        assert np.array_equal(g.view(np.uint32), want.view(np.uint32))
