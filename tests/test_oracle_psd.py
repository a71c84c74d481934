"""Pin the CPU oracle against the reference's own tests for the PSD path and
against an independent float64 numpy model.  CPU only."""
import numpy as np
import pytest

from conftest import uniform_noise
from oracle import model_f64 as m64


def test_hbf_passband_constant(oracle):
    # reference src/psd.rs:601: assert_eq!(idsp::hbf::HBF_PASSBAND, 0.4)
    assert oracle.lib().orc_hbf_passband() == np.float32(0.4)


@pytest.mark.parametrize("n", [16, 64, 512, 4096, 8192])
def test_fft_matches_dft_definition(oracle, n):
    rng = np.random.default_rng(n)
    c = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    got = oracle.fft_forward(c)
    want = np.fft.fft(c.astype(np.complex128))
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 3e-6


def test_window_matches_reference_formula(oracle):
    # src/psd.rs:42-55: periodic Hann, win[0] = 0, win[N/2] = 1, power 0.25, nenbw 1.5, overlap N/2
    for n in (16, 512, 4096):
        w, power, nenbw, ov = oracle.window(n, oracle.WINDOW_HANN)
        assert w[0] == 0.0 and abs(w[n // 2] - 1.0) < 1e-6 and w[1] != w[n - 1] or n == 16
        assert (power, nenbw, ov) == (0.25, 1.5, n // 2)
        np.testing.assert_allclose(w, np.sin(np.pi * np.arange(n) / n) ** 2, atol=2e-7)
        # exact mean-square and nenbw of the periodic window, independent of N
        assert abs(np.mean(w.astype(np.float64) ** 2) / np.mean(w.astype(np.float64)) ** 2 - 1.5) < 1e-5
        w, power, nenbw, ov = oracle.window(n, oracle.WINDOW_RECT)
        assert np.all(w == 1.0) and (power, nenbw, ov) == (1.0, 1.0, 0)


@pytest.mark.parametrize("det", [0, 1, 2, 3])
def test_detrend_apply(oracle, det):
    n = 512
    x = uniform_noise(n, 5) + np.float32(3.0) + np.linspace(0, 2, n, dtype=np.float32)
    w, *_ = oracle.window(n, oracle.WINDOW_HANN)
    got = oracle.detrend_apply(det, x, w)
    want = m64.detrend(x.astype(np.float64), det) * w
    assert np.all(got.imag == 0)
    # Span accumulates `offset += slope` sequentially in f32 (src/psd.rs:100): ~N*eps*|offset| drift
    np.testing.assert_allclose(got.real, want, atol=1e-4 if det == 2 else 2e-5)
    with pytest.raises(NotImplementedError):  # src/psd.rs:110 unimplemented!()
        oracle.detrend_apply(oracle.DETREND_LINEAR, x, w)


@pytest.mark.parametrize("n", [16, 512, 4096])
def test_known_answer_hann_ones(oracle, n):
    # The reference's commented-out exact test (src/psd.rs:562-597): Hann, x = ones -> PSD
    # [16/3, 4/3, 0] at N=4.  Generalised: [4N/3, N/3, 0, ...] for any N.
    s = oracle.Stage(n, oracle.WINDOW_HANN)
    s.process(np.ones(n, np.float32))
    assert s.count() == 1
    p = s.spectrum() / s.gain()
    assert abs(p[0] - 4 * n / 3) < 1e-5 * n and abs(p[1] - n / 3) < 1e-5 * n
    assert np.all(np.abs(p[2:]) < 1e-6 * n)


@pytest.mark.parametrize("preset", [0, 1])
def test_reference_statistical_test(oracle, preset):
    """Seeded re-run of the reference's live unit test, src/psd.rs:599-644."""
    x = uniform_noise(1 << 16, 0x7654321)
    xm = x.astype(np.float64).sum() / x.size
    assert abs(xm) < 10.0 / np.sqrt(x.size)
    xv = (x.astype(np.float64) ** 2).sum() / x.size
    assert abs(xv - 1.0) < 10.0 / np.sqrt(x.size)

    n = 1 << 9
    s = oracle.Stage(n, oracle.WINDOW_HANN, preset)
    y = s.process(x)
    # src/psd.rs:622
    assert y.size == (x.size >> 3) - oracle.lib().orc_hbf_response_length(preset)
    p = s.spectrum() / s.gain()
    assert s.count() == 255
    assert np.all(np.abs(p * 0.5 - 1.0) < 10.0 / np.sqrt(s.count()))  # src/psd.rs:626-632

    d = oracle.Cascade(n, preset)
    d.process(x)
    p, b = d.psd()
    assert len(b) >= 2
    for bi in b:  # src/psd.rs:637-643
        seg = p[bi.start:bi.start + bi.bins_end - bi.bins_start]
        if bi.include and bi.count:
            assert np.all(np.abs(seg * 0.5 - 1.0) < 10.0 / np.sqrt(bi.count))
            # continuity across stage breaks (needs the decimator's DC gain of 8 per stage)
            assert abs(np.mean(seg) * 0.5 - 1.0) < 5.0 / np.sqrt(bi.count * seg.size)
    f = oracle.break_frequencies(b)
    assert f.size == p.size and f[0] == 0.0 and f[-1] == 0.5  # src/psd.rs:324-325


@pytest.mark.parametrize("preset", [0, 1])
def test_hbf_matches_f64_model_and_spec(oracle, preset):
    x = uniform_noise(8 * 4096, 11)
    h = oracle.Hbf8(preset)
    # block-size independence (stateful across calls, src/psd.rs:246-253)
    y = np.concatenate([h.block(x[:8 * 100]), h.block(x[8 * 100:8 * 101]), h.block(x[8 * 101:])])
    want = m64.hbf8(x, preset)
    np.testing.assert_allclose(y, want, atol=3e-6)
    # DC gain 2 per half-band stage (8 per cascade), passband flat, alias band rejected
    for s in range(3):
        hh = m64.hbf_impulse(m64.TAPS[preset][s])
        assert abs(hh.sum() - 2.0) < 2e-4
    r = oracle.lib().orc_hbf_response_length(preset)
    # after R outputs the zero-state transient of a constant input has settled
    c = oracle.Hbf8(preset).block(np.ones(8 * 400, np.float32))
    assert np.all(np.abs(c[r:] - 8.0) < 1e-3)
    # a tone in the alias band of the output passband is suppressed by >= 90 dB
    t = np.arange(8 * 4096)
    tone = np.cos(2 * np.pi * (1.0 / 8 - 0.03) * t).astype(np.float32)  # aliases to 0.24 of out rate
    out = oracle.Hbf8(preset).block(tone)[r:]
    assert 20 * np.log10(np.sqrt(np.mean(out.astype(np.float64) ** 2)) / (8 * np.sqrt(0.5))) < (-90 if preset == 0 else -120)


@pytest.mark.parametrize("n,det,preset", [(512, 0, 1), (512, 3, 0), (4096, 1, 1), (64, 2, 1), (1024, 0, 1)])
def test_cascade_matches_f64_model(oracle, n, det, preset):
    x = uniform_noise(300 * n + 77, 3) + np.float32(0.25)
    c = oracle.Cascade(n, preset)
    c.set_detrend(det)
    # ragged feeding must not change anything
    rng = np.random.default_rng(1)
    pos = 0
    while pos < x.size:
        k = int(rng.integers(0, 5 * n))
        c.process(x[pos:pos + k])
        pos += k
    ref = m64.cascade(x, n, det, preset=preset)
    assert c.num_stages() == len(ref)
    for i, st in enumerate(ref):
        assert c.stage_count(i) == st["count"]
        got = c.stage_spectrum(i).astype(np.float64)
        scale = st["spectrum"].max() if st["count"] else 1.0
        np.testing.assert_allclose(got, st["spectrum"], atol=2e-5 * scale)
    p, b = c.psd()
    p64, b64 = m64.merge(ref, n)
    assert [bi.as_tuple() for bi in b] == [
        (d["start"], d["include"], d["count"], d["avg"], d["bins"][0], d["bins"][1], d["fft_size"],
         d["decimation"], d["pending"], d["processed"]) for d in b64]
    np.testing.assert_allclose(p, p64, rtol=1e-4, atol=1e-6)


def test_cascade_ewma_and_option_changes(oracle):
    n = 512
    x = uniform_noise(200 * n, 9)
    c = oracle.Cascade(n, 1)
    c.set_avg(limit=7, count=2 ** 32 - 2)  # like the psd binary: limit = avg_max - 1
    c.process(x)
    ref = m64.cascade(x, n, 0, avg_limit=7, avg_count=2 ** 32 - 2)
    for i, st in enumerate(ref):
        assert c.stage_count(i) == st["count"]
        np.testing.assert_allclose(c.stage_spectrum(i), st["spectrum"], rtol=2e-4, atol=1e-3)
    p, b = c.psd()
    assert b[-1].avg == 7 and b[-1].count == 8
    # stage counts of the effective (EWMA) average never exceed avg + 1
    assert all(bi.count <= bi.avg + 1 for bi in b)


def test_merge_options_and_breaks(oracle):
    n = 512
    c = oracle.Cascade(n, 1)
    c.process(uniform_noise(70 * 8 * n, 4))
    p, b = c.psd()
    # floor(2N/5) = 204, ceil(204/8) = 26
    assert [(bi.bins_start, bi.bins_end) for bi in b][-1] == (26, 257)
    assert b[0].bins_start == 0
    assert all(bi.bins_end == 204 for bi in b[:-1] if bi.include)
    assert [bi.decimation for bi in b] == [8 ** i for i in reversed(range(len(b)))]
    pk, bk = c.psd(keep_overlap=True, min_count=0, keep_transition_band=True)
    assert all((bi.bins_start, bi.bins_end) == (0, 257) for bi in bk)
    assert pk.size == 257 * len(bk)
    # min_count larger than the deepest stages' counts drops them and restarts the low edge
    pm, bm = c.psd(min_count=20)
    first_inc = next(bi for bi in bm if bi.include)
    assert first_inc.bins_start == 0 and not bm[0].include
    assert abs(c.rbw() - 8 / (512 * 0.4)) < 1e-6  # src/psd.rs:427-429


def test_empty_and_short_inputs(oracle):
    c = oracle.Cascade(512, 1)
    c.process(np.zeros(0, np.float32))
    assert c.num_stages() == 0 and c.psd()[0].size == 0
    c.process(np.ones(100, np.float32))
    p, b = c.psd()
    assert c.num_stages() == 1 and p.size == 0 and b[0].pending == 100 and not b[0].include
    s = oracle.Stage(512)
    assert s.process(np.ones(511, np.float32)).size == 0 and s.buf().size == 511 and s.count() == 0
    s.process(np.ones(1, np.float32))
    assert s.count() == 1 and s.buf().size == 256


@pytest.mark.parametrize("preset", [0, 1])
def test_first_segment_inside_a_full_chunk(oracle, preset):
    """A short first call followed by a full 8N chunk makes stage 0 emit N/8 - drain + 15 N/16 > N items
    in one chunk: the reference's [f32; N] ping-pong arrays (psd.rs:457) are too small for that and its
    slice at psd.rs:253 panics; the restatement continues the arithmetic, and must stay call-size invariant."""
    n = 512
    x = uniform_noise(n - 1 + 8 * n + 3 * n, 21)
    a = oracle.Cascade(n, preset)
    a.process(x[:n - 1])
    a.process(x[n - 1:])
    b = oracle.Cascade(n, preset)
    for i in range(0, x.size, 300):
        b.process(x[i:i + 300])
    pa, ba = a.psd(True, 0, True)
    pb, bb = b.psd(True, 0, True)
    assert [k.as_tuple() for k in ba] == [k.as_tuple() for k in bb]
    assert np.array_equal(pa, pb)


def test_var_kat(oracle):
    # reference src/var.rs:52-60
    v = oracle.var_eval([1000.0, 100.0, 1.2, 3.4, 5.6], [0.0, 1.0, 3.0, 6.0, 9.0], 2.7)
    assert abs(0.13478442 - v) < 1e-6
