"""Device-resident synthetic sources (source.rs:104-134) against the CPU restatement.  Needs a B200.

White and differentiated noise and the MASH-1-1-1 marker are bit-exact.  Integrated noise is not
comparable bit for bit: the reference accumulates sequentially in f32 (source.rs:113), the device
evaluates the same recurrence as a scan carried in f64, which is the more accurate of the two; it is
therefore checked against the f64 evaluation of the recurrence to f32 rounding, and against the f32
restatement to the drift that restatement itself shows against f64."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CHUNKS = (1, 3, 4092, 4096, 65537, 100_000, 830_271)  # odd cuts: the stream must not depend on them


@pytest.fixture(scope="module")
def sp():
    import torch
    assert torch.cuda.is_available()
    import stabilizer_stream_b200 as m
    return m


def gpu_stream(src, chunks):
    import torch
    out = [src.get(n) for n in chunks]
    torch.cuda.synchronize()
    return np.concatenate([o.cpu().numpy() for o in out])


@pytest.mark.parametrize("param", [0, 1, 2, 3, 4, 5, 8])
def test_white_and_differentiated_noise_bit_exact(sp, oracle, param):
    n = sum(CHUNKS)
    want = oracle.Source(oracle.SOURCE_NOISE, param).get(n)
    got = gpu_stream(sp.Source.noise(param), CHUNKS)
    assert np.array_equal(got, want)
    one = gpu_stream(sp.Source.noise(param), (n,))
    assert np.array_equal(one, want)


def integrate_f64(w, order):
    x = w.astype(np.float64)
    for _ in range(order):
        c = np.cumsum(x)
        x = np.concatenate([[0.0], c[:-1]])  # (x, s) = (s, x + s): the old state is the output
    return x


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_integrated_noise_matches_the_f64_recurrence(sp, oracle, order):
    n = sum(CHUNKS)
    w = oracle.Source(oracle.SOURCE_NOISE, 0).get(n)
    want = integrate_f64(w, order)
    got = gpu_stream(sp.Source.noise(-order), CHUNKS).astype(np.float64)
    scale = np.maximum.accumulate(np.abs(want)) + 1.0
    # one f32 rounding of the output (2^-24 relative) plus f64 scan noise
    assert np.max(np.abs(got - want) / scale) <= 1e-7
    one = gpu_stream(sp.Source.noise(-order), (n,)).astype(np.float64)
    assert np.max(np.abs(one - want) / scale) <= 1e-7
    if order == 1:
        # the sequential f32 restatement drifts from f64 by its own accumulated rounding; the
        # device stream stays within that drift of it
        ref32 = oracle.Source(oracle.SOURCE_NOISE, -1).get(n).astype(np.float64)
        drift = np.max(np.abs(ref32 - want) / scale)
        assert np.max(np.abs(got - ref32) / scale) <= drift + 1e-7
        assert drift < 1e-3


@pytest.mark.parametrize("ftw", [0x1000000, 0x12345678, 1, 0xfffffff1])
def test_dsm_marker_bit_exact(sp, oracle, ftw):
    n = sum(CHUNKS)
    want = oracle.Source(oracle.SOURCE_DSM, ftw).get(n)
    got = gpu_stream(sp.Source.dsm(ftw), CHUNKS)
    assert np.array_equal(got, want)
    one = gpu_stream(sp.Source.dsm(ftw), (n,))
    assert np.array_equal(one, want)


def test_seed_position_reset(sp, oracle):
    s = sp.Source.noise(0, seed=99)
    a = gpu_stream(s, (1000,))
    assert s.position() == 1000
    assert np.array_equal(a, oracle.Source(oracle.SOURCE_NOISE, 0, seed=99).get(1000))
    s.reset()
    assert s.position() == 0
    assert np.array_equal(gpu_stream(s, (1000,)), a)
    for bad in (-5, 9):
        with pytest.raises(Exception):
            sp.Source.noise(bad)


@pytest.mark.parametrize("param", [0, 1, -1])
def test_process_source_equals_process_of_the_generated_stream(sp, oracle, param):
    import torch
    n = 3_000_001
    a = sp.PsdCascade(512)
    a.set_detrend(sp.Detrend.MEAN)
    a.process_source(sp.Source.noise(param), n)
    b = sp.PsdCascade(512)
    b.set_detrend(sp.Detrend.MEAN)
    x = sp.Source.noise(param).get(n)
    b.process(x)
    torch.cuda.synchronize()
    pa, ba = a.psd()
    pb, bb = b.psd()
    assert [(k.count, k.pending, k.processed) for k in ba] == [(k.count, k.pending, k.processed) for k in bb]
    np.testing.assert_allclose(pa, pb, rtol=1e-5)
    # and the CPU cascade fed the CPU restatement of the stream agrees within the PSD tolerance
    if param >= 0:
        o = oracle.Cascade(512)
        o.set_detrend(oracle.DETREND_MEAN)
        o.process(oracle.Source(oracle.SOURCE_NOISE, param).get(n))
        po, _ = o.psd()
        # bins far below the stream's level sit on the f32 rounding floor of the decimators, where the
        # two implementations legitimately differ (differentiated noise vanishes towards DC)
        po = np.asarray(po)
        inc = po > 1e-4 * np.median(po)
        np.testing.assert_allclose(np.asarray(pa)[inc][4:], np.asarray(po)[inc][4:], rtol=1e-3)


@pytest.mark.parametrize("param", [1, 2, 3, -1])
def test_power_law_shape(sp, param):
    """each differentiator / integrator multiplies the PSD by (2 sin(pi f))^(+-2); white level is 2
    in the reference's normalisation (0.5 * p ~ 1, psd.rs:629).  (Two or more integrators are not
    testable this way: the f32 samples of a 2^24-step double random walk are quantised far above the
    f^-4 law's high-frequency level, in the reference as much as here.)"""
    n = 1 << 24
    g = sp.Psd(1024, sp.Window.HANN)
    g.set_detrend(sp.Detrend.MEAN)
    x = sp.Source.noise(param).get(n)
    g.process(x)
    p = np.asarray(g.spectrum(), np.float64) / g.gain()
    f = np.arange(p.size) / 1024.0
    k = np.arange(32, 480)  # away from DC leakage and Nyquist
    model = 2.0 * (2.0 * np.sin(np.pi * f[k])) ** (2 * param)
    ratio = p[k] / model
    assert abs(np.mean(ratio) - 1.0) < 0.02
    assert np.max(np.abs(ratio - 1.0)) < 10.0 / np.sqrt(g.count())
