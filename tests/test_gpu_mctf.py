"""GPU: motion estimation of the GOP-based temporal filter (vtmme_mctf_me) against the oracle — which is pinned on the
reference's own EncTemporalFilter::motionEstimation — and against committed vectors of the reference."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bindings as B  # noqa: E402
from tests.helpers import pad_plane  # noqa: E402


@pytest.fixture(scope="module")
def ms():
    import vtm_b200
    m = vtm_b200.MotionSearch(0)
    yield m
    m.close()


def oracle_mctf(L, cur, ref, bd):
    h, w = cur.shape
    curp, refp = pad_plane(cur, 128), pad_plane(ref, 128)
    stride = curp.shape[1]
    off = 128 * stride + 128
    out = np.zeros((h // 4, w // 4, 3), np.int32)
    L.vo_mctf_me(B.ptr(curp, off), stride, B.ptr(refp, off), stride, w, h, bd, C.c_void_p(out.ctypes.data))
    return out


@pytest.mark.parametrize("w,h,bd", [(208, 120, 10), (176, 144, 8), (64, 48, 10), (416, 240, 10)])
def test_mctf_me_matches_oracle(ms, oracle_lib, w, h, bd):
    """Whole pyramid, every vector and error; three references against one original in one call; sizes that are not
    multiples of 16 / 32 (partial last blocks are skipped like the reference skips them)."""
    from vtm_b200.synth import make_pair
    pairs = [make_pair(400 + w + k, w, h, max_global=9, max_local=14, n_rects=3, sigma=5.0, bit_depth=bd) for k in range(3)]
    cur = pairs[0][0]
    ms.upload_picture(70, cur)
    for k in range(3):
        ms.upload_picture(71 + k, pairs[k][1] if k == 0 else pairs[k][0])   # reference frames: one true, two unrelated
    got = ms.mctf_me([70, 70, 70], [71, 72, 73], w, h, bd)
    refs = [pairs[0][1], pairs[1][0], pairs[2][0]]
    for k in range(3):
        want = oracle_mctf(oracle_lib, cur, refs[k], bd)
        assert np.array_equal(got[k], want), (k, np.argwhere(got[k] != want)[:4])
    assert (got[0][:, :, 2] != np.iinfo(np.int32).max).sum() == ((h - 1) // 8) * ((w - 1) // 8)


def test_mctf_me_golden(ms):
    """Committed vectors of the reference itself (tests/golden/mctf_golden.npz, make_golden_mctf.py)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mctf_golden.npz"))
    cur, ref = np.ascontiguousarray(g["cur"]), np.ascontiguousarray(g["ref"])
    h, w = cur.shape
    ms.upload_picture(74, cur)
    ms.upload_picture(75, ref)
    assert np.array_equal(ms.mctf_me([74], [75], w, h, 10)[0], g["mv"])


@pytest.mark.parametrize("w,h,bd", [(208, 120, 10), (100, 68, 8)])
def test_mctf_apply_motion(ms, oracle_lib, w, h, bd):
    """applyMotion for luma and 4:2:0 chroma with the GPU's own vectors and with random 1/16-sample vectors; a size with
    partial blocks at the right / bottom edge."""
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(600 + w)
    cur, ref, _ = make_pair(320 + w, w, h, max_global=9, max_local=14, n_rects=3, sigma=5.0, bit_depth=bd)
    ref = np.ascontiguousarray(ref)
    chroma = np.ascontiguousarray(rng.integers(0, 1 << bd, (h // 2, w // 2), dtype=np.int16))
    ms.upload_picture(76, np.ascontiguousarray(cur))
    ms.upload_picture(77, ref)
    ms.upload_picture(78, chroma)
    mv = ms.mctf_me([76], [77], w, h, bd)[0]
    refp, chp = pad_plane(ref, 128), pad_plane(chroma, 64)
    for variant in range(2):
        if variant:
            mv = mv.copy()
            mv[:, :, 0] = rng.integers(-300, 301, mv.shape[:2])
            mv[:, :, 1] = rng.integers(-300, 301, mv.shape[:2])
        want_y, want_c = np.zeros((h, w), np.int16), np.zeros((h // 2, w // 2), np.int16)
        oracle_lib.vo_mctf_apply_motion(B.ptr(refp, 128 * refp.shape[1] + 128), refp.shape[1], w, h, 0, 0,
                                        C.c_void_p(mv.ctypes.data), mv.shape[1], bd, B.ptr(want_y), w)
        oracle_lib.vo_mctf_apply_motion(B.ptr(chp, 64 * chp.shape[1] + 64), chp.shape[1], w // 2, h // 2, 1, 1,
                                        C.c_void_p(mv.ctypes.data), mv.shape[1], bd, B.ptr(want_c), w // 2)
        assert np.array_equal(ms.mctf_apply_motion(77, w, h, mv, 0, 0, bd), want_y), variant
        assert np.array_equal(ms.mctf_apply_motion(78, w // 2, h // 2, mv, 1, 1, bd), want_c), variant


@pytest.mark.parametrize("num_refs,bd,qp,strength", [(4, 10, 32, 0.95), (2, 10, 27, 1.5), (3, 8, 37, 0.95), (1, 10, 22, 1.5)])
def test_mctf_bilateral(ms, oracle_lib, num_refs, bd, qp, strength):
    """The whole filter of one picture on the GPU — motion estimation against every neighbour, applyMotion, bilateral
    weighting (luma and 4:2:0 chroma) — against the oracle's vo_mctf_bilateral, which is pinned on the reference's own
    EncTemporalFilter::bilateralFilter (tests/test_oracle_vs_ref.py).  The weight tables come from the host's exp()."""
    from vtm_b200.synth import make_pair
    w, h = 208, 120
    rng = np.random.default_rng(1100 + num_refs)
    org, _, _ = make_pair(160, w, h, max_global=5, max_local=8, n_rects=3, sigma=5.0, bit_depth=bd)
    org = np.ascontiguousarray(org)
    org_c = np.ascontiguousarray(rng.integers(0, 1 << bd, (h // 2, w // 2), dtype=np.int16))
    offsets = [[-1], [-1, 1], [-2, -1, 1], [-2, -1, 1, 2]][num_refs - 1]
    ms.upload_picture(500, org)
    ms.upload_picture(501, org_c)
    corr_y, corr_c = [], []
    for k in range(num_refs):
        r = np.clip(np.roll(org.astype(np.int32), (k + 1, -2 * k - 1), (0, 1)) + np.rint(rng.normal(0, 3 + 2 * k, org.shape)).astype(np.int32),
                    0, (1 << bd) - 1).astype(np.int16)
        c = np.clip(org_c.astype(np.int32) + np.rint(rng.normal(0, 6, org_c.shape)).astype(np.int32), 0, (1 << bd) - 1).astype(np.int16)
        ms.upload_picture(510 + k, np.ascontiguousarray(r))
        ms.upload_picture(520 + k, np.ascontiguousarray(c))
        mv = ms.mctf_me([500], [510 + k], w, h, bd)[0]
        corr_y.append(ms.mctf_apply_motion(510 + k, w, h, mv, 0, 0, bd))
        corr_c.append(ms.mctf_apply_motion(520 + k, w // 2, h // 2, mv, 1, 1, bd))
        ms.upload_picture(530 + k, corr_y[-1])
        ms.upload_picture(540 + k, corr_c[-1])
    offs = (C.c_int * num_refs)(*offsets)
    for comp, (o, oid, corr, base, cw, ch) in enumerate([(org, 500, corr_y, 530, w, h), (org_c, 501, corr_c, 540, w // 2, h // 2)]):
        tables = []
        for k in range(num_refs):
            tb = np.zeros(1 << bd, np.float64)
            oracle_lib.vo_mctf_bilateral_weights(comp, qp, strength, bd, num_refs, min(1, abs(offsets[k]) - 1), C.c_void_p(tb.ctypes.data))
            tables.append(tb)
        got = ms.mctf_bilateral(oid, [base + k for k in range(num_refs)], np.stack(tables), cw, ch, bd)
        cptrs = (C.c_void_p * num_refs)(*[a.ctypes.data for a in corr])
        want = np.zeros((ch, cw), np.int16)
        oracle_lib.vo_mctf_bilateral(B.ptr(o), cw, cptrs, cw, offs, num_refs, cw, ch, comp, qp, strength, bd, B.ptr(want), cw)
        assert np.array_equal(got, want), (comp, np.argwhere(got != want)[:5])
        assert (got != o).mean() > 0.2
