"""Trace::plot / struct Trapezoidal (reference src/bin/psd.rs:96-157): the library's host helper
sspsd_trace_plot against a hand-derived known answer, the oracle restatement and an f64 trapezoid."""
import numpy as np

from oracle import binding as orc
from stabilizer_stream_b200 import Trace


def test_known_answer():
    # trapezoids between (0,2) (0.1,2) (0.2,4) (0.4,4): 0.2 + 0.3 + 0.8; the push of the first point adds 0
    f = np.array([0.0, 0.1, 0.2, 0.4], np.float32)
    p = np.array([2.0, 2.0, 4.0, 4.0], np.float32)
    integ, xy = Trace("t", [], p, f).plot()
    assert abs(integ - np.sqrt(1.3)) < 1e-6
    assert xy.shape == (3, 2)                     # DC is not `is_normal()`: skipped (bin/psd.rs:140)
    assert np.allclose(xy[:, 0], np.log10(f[1:].astype(np.float64)), atol=1e-6)
    assert np.allclose(xy[:, 1], 10 * np.log10(p[1:].astype(np.float64)), atol=1e-5)
    # integrate = true plots sqrt of the running integral from DC (bin/psd.rs:145-146)
    _, xyi = Trace("t", [], p, f).plot(integrate=True)
    assert np.allclose(xyi[:, 1], np.sqrt([0.2, 0.5, 1.3]), atol=1e-6)
    # band limits act on fs * f, inclusive at both ends (bin/psd.rs:136)
    integ2, xy2 = Trace("t", [], p, f).plot(fs=10.0, integral_start=2.0, integral_end=4.0)
    assert abs(integ2 - np.sqrt(0.3 + 0.8)) < 1e-6
    assert np.allclose(xy2[:, 0], np.log10(f[1:].astype(np.float64)) + 1.0, atol=1e-6)
    assert np.allclose(xy2[:, 1], 10 * (np.log10(p[1:].astype(np.float64)) - 1.0), atol=1e-5)


def test_matches_oracle_bitwise_on_a_merged_spectrum():
    rng = np.random.default_rng(5)
    x = ((rng.random(1 << 20, dtype=np.float32) - np.float32(0.5)) * np.float32(12 ** 0.5)).astype(np.float32)
    c = orc.Cascade(512, orc.HBF_140)
    c.process(x)
    p, b = c.psd()
    f = orc.break_frequencies(b)
    for kw in (dict(), dict(fs=200e6, integral_start=1e3, integral_end=1e8), dict(integrate=True, fs=3.0)):
        gi, gxy = Trace("t", [], p, f).plot(**kw)
        oi, oxy = orc.trace_plot(p, f, **kw)
        assert np.float32(gi).view(np.uint32) == np.float32(oi).view(np.uint32)
        assert np.array_equal(gxy.view(np.uint64), oxy.view(np.uint64))
    # white noise of unit variance: the one-sided PSD integrates to ~1 over the whole band
    gi, _ = Trace("t", [], p, f).plot()
    ref = np.sqrt(np.trapezoid(p.astype(np.float64), f.astype(np.float64)))
    assert abs(gi - ref) < 1e-4 * ref and abs(gi - 1.0) < 0.05


def test_empty_and_capacity():
    integ, xy = Trace("t", [], np.zeros(0, np.float32), np.zeros(0, np.float32)).plot()
    assert integ == 0.0 and xy.shape == (0, 2)
