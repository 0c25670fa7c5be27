"""The multiply-high division of the table-level filter kernels (vtm_b200/csrc/misc_kernels.cu, struct FastDiv), restated in
Python: q = (x * m) >> 32 with m = floor((2^32 - 1) / d) + 1 equals x // d for every x the host-side bound admits
(m * d = 2^32 + e; exact while x * e < 2^32; always for powers of two)."""
import numpy as np


def fastdiv_params(d):
    m = ((0xFFFFFFFF // d) + 1) & 0xFFFFFFFF
    return m


def exact_below(d, max_x):
    if d == 1:
        return max_x <= 0xFFFFFFFF
    m = fastdiv_params(d)
    e = m * d - (1 << 32)
    return max_x <= 0xFFFFFFFF and (e == 0 or max_x * e < (1 << 32))


def fastdiv(x, d):
    return x if d == 1 else (x * fastdiv_params(d)) >> 32


def test_fastdiv_exact_within_the_host_bound():
    rng = np.random.default_rng(11)
    divisors = [1, 2, 3, 4, 5, 6, 7, 8, 12, 16, 24, 32, 48, 64, 96, 128] + [int(v) for v in rng.integers(2, 5000, 60)]
    for d in divisors:
        m = fastdiv_params(d)
        e = (m * d - (1 << 32)) if d > 1 else 0
        assert 0 <= e <= d
        # the largest dividend the bound admits (capped at 2^31: the kernels' unit counts)
        hi = (1 << 31) - 1
        if e:
            hi = min(hi, ((1 << 32) - 1) // e)
        assert exact_below(d, hi)
        xs = np.concatenate([rng.integers(0, hi + 1, 4000), np.array([0, 1, d - 1, d, d + 1, hi - 1, hi]),
                             (np.arange(1, 400) * d - 1).clip(0, hi), (np.arange(0, 400) * d).clip(0, hi)])
        for x in xs:
            x = int(x)
            assert fastdiv(x, d) == x // d, (d, x)


def test_fastdiv_bound_is_not_vacuous():
    # beyond the bound the quotient does go wrong for some x: the host check is what makes the device code safe
    d = 6
    m = fastdiv_params(d)
    e = m * d - (1 << 32)
    bad = [x for x in range(((1 << 32) // e) - 20, ((1 << 32) // e) + 200000) if ((x * m) >> 32) != x // d]
    assert e > 0 and bad and min(bad) * e >= (1 << 32)


def test_cross_half_stage_identity():
    """|a + b| + |a - b| == 2 max(|a|, |b|): what lets the two threads of a 16x8 / 8x16 SATD tile swap transformed coefficients
    instead of running the last butterfly stage (satd_tile_thread_kernel)."""
    rng = np.random.default_rng(12)
    a = rng.integers(-(1 << 23), 1 << 23, 100000)
    b = rng.integers(-(1 << 23), 1 << 23, 100000)
    assert np.array_equal(np.abs(a + b) + np.abs(a - b), 2 * np.maximum(np.abs(a), np.abs(b)))
