"""CPU tests of the synthetic-source restatement (oracle/, source.rs:66-73, 104-134)."""
import numpy as np
import pytest

from oracle import binding as B

# Random123 known-answer vectors for Philox4x32-10 (counter, key, output)
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_known_answers(ctr, key, want):
    assert B.philox4x32_10(ctr, key) == want


def test_white_noise_moments_and_range():
    x = B.Source(B.SOURCE_NOISE, 0).get(1 << 20)
    assert abs(float(x.mean())) < 5e-3
    assert abs(float(x.std()) - 1.0) < 5e-3
    # Open01: strictly inside (-sqrt(3), sqrt(3))
    assert float(np.abs(x).max()) < np.sqrt(3.0)


@pytest.mark.parametrize("param", [0, 1, -1, 3, -2])
def test_stream_does_not_depend_on_call_sizes(param):
    a = B.Source(B.SOURCE_NOISE, param).get(10_000)
    s = B.Source(B.SOURCE_NOISE, param)
    b = np.concatenate([s.get(n) for n in (1, 4095, 4096, 1808)])
    assert np.array_equal(a, b)


def test_differentiators_and_integrators_match_the_fold():
    w = B.Source(B.SOURCE_NOISE, 0).get(5000)
    # noise > 0: (x, s) = (x - s, x), source.rs:112
    d1 = B.Source(B.SOURCE_NOISE, 1).get(5000)
    assert np.array_equal(d1, w - np.concatenate([[np.float32(0)], w[:-1]]))
    d2 = B.Source(B.SOURCE_NOISE, 2).get(5000)
    assert np.array_equal(d2, d1 - np.concatenate([[np.float32(0)], d1[:-1]]))
    # noise < 0: (x, s) = (s, x + s): delayed running sum in f32
    i1 = B.Source(B.SOURCE_NOISE, -1).get(5000)
    c = np.cumsum(w, dtype=np.float32)
    assert i1[0] == 0 and np.array_equal(i1[1:], c[:-1])


def test_seed_selects_the_stream():
    a = B.Source(B.SOURCE_NOISE, 0, seed=1).get(100)
    b = B.Source(B.SOURCE_NOISE, 0, seed=2).get(100)
    assert not np.array_equal(a, b)


@pytest.mark.parametrize("ftw", [0x1000000, 0x12345678, 1])
def test_dsm_is_a_third_order_modulator(ftw):
    n = 1 << 16
    y = B.Source(B.SOURCE_DSM, ftw).get(n).astype(np.float64) + 0.5
    assert set(np.unique(y)).issubset(set(range(-3, 5)))  # idsp Dsm<3> output range 1-(1<<2) ..= 1<<2
    x = np.array([B.dsm_input(i, ftw) for i in range(n)], dtype=np.float64) / 2.0 ** 32
    # MASH-1-1-1: y = x + (1 - z^-1)^3 e3 with e3 the (bounded) residue of the last accumulator,
    # so integrating the error three times must stay within one unit
    e = y - x
    for _ in range(3):
        e = np.cumsum(e)
    assert np.abs(e).max() <= 1.0 + 1e-6
