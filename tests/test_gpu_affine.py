"""GPU: the affine-ME primitives (AffineGradientSearch's dispatch-table entries) through the C ABI against the oracle, which
tests/test_oracle_vs_ref.py pins on the reference's own (SIMD) entries.  Bit-exact: every derivative, every int64 sum."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bindings as B  # noqa: E402


@pytest.fixture(scope="module")
def ms():
    import vtm_b200
    m = vtm_b200.MotionSearch(0)
    yield m
    m.close()


def _oracle(L, org, pred, six):
    h, w = pred.shape
    d = [np.zeros((h, w), np.int32) for _ in range(2)]
    for v in (0, 1):
        L.vo_affine_sobel(v, B.ptr(pred), pred.strides[0] // 2, C.c_void_p(d[v].ctypes.data), w, w, h)
    res = np.ascontiguousarray((org.astype(np.int32) - pred.astype(np.int32)).astype(np.int16))
    c = np.zeros((7, 7), np.int64)
    L.vo_affine_equal_coeff(B.ptr(res), w, C.c_void_p(d[0].ctypes.data), C.c_void_p(d[1].ctypes.data), w, C.c_void_p(c.ctypes.data), w, h, six)
    return d, res, c


@pytest.mark.parametrize("six", [0, 1])
def test_affine_primitives(ms, oracle_lib, six):
    """Table entries one block at a time (strided views, accumulation into a non-zero matrix) and the batched gradient step
    (mixed sizes in one launch, bi-predictive residual range)."""
    rng = np.random.default_rng(1400 + six)
    blocks, want = [], []
    for (w, h) in [(16, 16), (32, 16), (16, 64), (64, 64), (128, 32), (128, 128), (8, 8), (4, 8)]:
        big = rng.integers(0, 1024, (h + 3, w + 11)).astype(np.int16)
        pred = big[2:2 + h, 5:5 + w]                                   # a view: row stride w + 11
        org = np.ascontiguousarray(np.clip(pred.astype(np.int32) + rng.integers(-300, 301, (h, w)), -1023, 2046).astype(np.int16))
        d, res, c = _oracle(oracle_lib, org, pred, six)
        for v in (0, 1):
            assert np.array_equal(ms.affine_sobel(v, pred), d[v]), (w, h, v)
        start = np.zeros((7, 7), np.int64)
        start[2, 1] = -777
        got = ms.affine_equal_coeff(res, d[0], d[1], six, start.copy())
        assert np.array_equal(got - start, c), (w, h)
        blocks.append((org, pred, six))
        want.append(c)
    got = ms.affine_gradient_step(blocks)
    for i in range(len(blocks)):
        assert np.array_equal(got[i], want[i]), (i, blocks[i][0].shape)
    assert np.abs(got).max() > 1 << 32
