"""CPU check of the arithmetic fact K1's `key_less` (csrc/halfplane_dev.cuh) rests on: the reference compares ROUNDED
distances sqrt(d2) (`ObstaclesUtils.py:91-94`); the kernel decides the same comparison from the squared distances and
only takes square roots inside a relative gap of 2^-48.  Outside that gap the two must agree for every pair."""
import numpy as np


def key_less_restated(a, b):
    """Mirror of the device function (EXACT = true) in numpy float64."""
    out = np.zeros(a.shape, bool)
    lt = a < b
    thr = b - np.ldexp(b, -48)                     # fma(-2^-48, b, b) up to one rounding; the margin absorbs it
    sure = lt & (a < thr)
    near = lt & ~sure
    out[sure] = True
    out[near] = np.sqrt(a[near]) < np.sqrt(b[near])
    return out


def test_squared_distance_comparison_equals_rounded_distance_comparison():
    rs = np.random.default_rng(0)
    b = np.concatenate([rs.uniform(1e-12, 50.0, 200000), np.ldexp(rs.uniform(1, 2, 50000), rs.integers(-60, 20, 50000))])
    # a at 0 .. 2^12 ulps below b, and exactly equal, and above
    k = rs.integers(-8, 1 << 12, b.shape)
    a = b.copy()
    for _ in range(1):
        a = (b.view(np.int64) - k).view(np.float64)
    want = np.sqrt(a) < np.sqrt(b)
    got = key_less_restated(a, b)
    assert np.array_equal(want, got)
    # the fast "sure" branch is only ever taken when the rounded roots really differ
    thr = b - np.ldexp(b, -48)
    sure = (a < b) & (a < thr)
    assert np.all(np.sqrt(a[sure]) < np.sqrt(b[sure]))
    # and it is the common case for anything but near ties: random pairs
    a2 = rs.uniform(1e-12, 50.0, b.shape)
    assert np.array_equal(np.sqrt(a2) < np.sqrt(b), key_less_restated(a2, b))
