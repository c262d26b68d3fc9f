"""fastdiv (csrc/dofs_common.cuh): x / d for a launch-constant divisor as one multiply-high and ONE correction step.
The kernels rely on floor(x * m / 2^32) being floor(x / d) or one less for every 0 <= x < 2^31 with
m = floor((2^32 - 1) / d); this checks that bound on the integers (the device code is the same three lines)."""
import numpy as np


def fastdiv(x, d, m):
    q = (x * m) >> 32          # __umulhi
    r = x - q * d
    fix = r >= d               # the single correction step
    return q + fix, r - fix * d


def test_fastdiv_is_exact_below_2_31():
    rng = np.random.default_rng(5)
    widths = [1, 2, 3, 5, 7, 64, 327, 328, 640, 1919, 1920, 3840, 4096, 65535, 65536, (1 << 24) - 1, (1 << 30) + 1]
    widths += [int(v) for v in rng.integers(1, 1 << 20, 40)]
    edge = np.array([0, 1, 2, (1 << 31) - 1, (1 << 31) - 2, (1 << 30), (1 << 30) - 1], dtype=np.uint64)
    for d in widths:
        m = np.uint64(0xFFFFFFFF // d)
        xs = np.concatenate([edge, rng.integers(0, 1 << 31, 20000).astype(np.uint64),
                             (np.arange(0, 2000, dtype=np.uint64) * np.uint64(d)) % np.uint64(1 << 31),
                             ((np.arange(1, 2000, dtype=np.uint64) * np.uint64(d)) - np.uint64(1)) % np.uint64(1 << 31)])
        q, r = fastdiv(xs, np.uint64(d), m)
        assert np.array_equal(q, xs // np.uint64(d)), d
        assert np.array_equal(r, xs % np.uint64(d)), d
        # and never more than one below before the correction
        q0 = (xs * m) >> np.uint64(32)
        assert ((xs // np.uint64(d)) - q0).max() <= 1, d
