"""GPU: decoder-side MV refinement (vtmme_dmvr_refine) through the C ABI against the oracle's restatement of the search of
InterPrediction::xProcessDMVR (pinned on the reference's own members) and against the committed fixtures the reference
produced (tests/golden/dmvr_golden.npz).  Bit-exact."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bindings as B  # noqa: E402
from tests.helpers import MARGIN, dmvr_cases, pad_plane  # noqa: E402


@pytest.fixture(scope="module")
def ms():
    import vtm_b200
    m = vtm_b200.MotionSearch(0)
    yield m
    m.close()


def oracle_dmvr(L, p0, p1, margin, blk, pic_w, pic_h, bd, max_cu=128):
    stride, off = p0.shape[1], margin * p0.shape[1] + margin
    out = np.zeros((len(blk), 4), np.int32)
    for i, b in enumerate(blk):
        L.vo_dmvr_block(B.ptr(p0, off), B.ptr(p1, off), stride, *[int(v) for v in b], pic_w, pic_h, max_cu, max_cu, bd,
                        C.c_void_p(out[i].ctypes.data))
    return out


def test_dmvr_golden(ms):
    from tests.test_golden import iter_dmvr
    for k, (p0, p1, m, blk, want) in enumerate(iter_dmvr()):
        ms.upload_picture(70, p0, m)
        ms.upload_picture(71, p1, m)
        got = ms.dmvr_refine(70, 71, blk, 10)
        assert np.array_equal(got, want), (k, np.argwhere(got != want)[:4])


@pytest.mark.parametrize("bd,max_cu", [(10, 128), (8, 128), (10, 64)])
def test_dmvr_matches_oracle(ms, oracle_lib, bd, max_cu):
    """All phases of the bilinear filter, early exits, integer and sub-sample outcomes, MVs clipped at the picture border,
    8 and 10 bit, both CTU sizes of the MV clip; the border of the second picture is produced on the device."""
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(950 + bd + max_cu)
    W, H = 256, 192
    moved = sub = skipped = 0
    for seed in range(3):
        r0, r1, _ = make_pair(130 + seed, W, H, max_global=[1, 5, 2][seed], max_local=3, n_rects=4, sigma=[2.0, 6.0, 3.0][seed],
                              bit_depth=bd)
        if seed == 2:
            r1 = r0
        p0, p1 = pad_plane(r0), pad_plane(r1)
        ms.upload_picture(72, p0, MARGIN)
        ms.upload_picture(73, np.ascontiguousarray(r1))          # no border: replicated on the device
        blk = dmvr_cases(rng, W, H, 600, same=seed == 2)
        got = ms.dmvr_refine(72, 73, blk, bd, max_cu)
        want = oracle_dmvr(oracle_lib, p0, p1, MARGIN, blk, W, H, bd, max_cu)
        assert np.array_equal(got, want), (seed, np.argwhere(got != want)[:4])
        moved += int((want[:, :2] != 0).any(axis=1).sum())
        sub += int((want[:, :2] % 16 != 0).any(axis=1).sum())
        skipped += int((want[:, 3] == 0).sum())
    assert moved > 200 and sub > 100 and skipped > 30


def test_dmvr_full_picture(ms, oracle_lib):
    """Every 16x16 sub-block of a 1080p picture in one call (8,040 blocks, one launch); a seeded sample against the
    oracle, the rest through properties: reading one picture through both lists with opposite MVs is an exact match
    (early exit, zero refinement), and the call is idempotent."""
    from vtm_b200.synth import make_pair
    W, H = 1920, 1080
    r0, r1, _ = make_pair(5, W, H, max_global=3, max_local=4)
    ms.upload_picture(74, np.ascontiguousarray(r0))
    ms.upload_picture(75, np.ascontiguousarray(r1))
    rng = np.random.default_rng(77)
    xs, ys = np.meshgrid(np.arange(0, W - 15, 16), np.arange(0, H - 15, 16))
    n = xs.size
    mv = rng.integers(-6 * 16, 6 * 16 + 1, (n, 2))
    blk = np.stack([xs.ravel(), ys.ravel(), np.full(n, 16), np.full(n, 16), mv[:, 0], mv[:, 1], -mv[:, 0] + rng.integers(-24, 25, n),
                    -mv[:, 1] + rng.integers(-24, 25, n)], axis=1).astype(np.int32)
    launches = ms.launches
    got = ms.dmvr_refine(74, 75, blk)
    assert ms.launches - launches == 1
    pick = rng.choice(n, 300, replace=False)
    want = oracle_dmvr(oracle_lib, pad_plane(r0), pad_plane(r1), MARGIN, blk[pick], W, H, 10)
    assert np.array_equal(got[pick], want)
    assert np.array_equal(ms.dmvr_refine(74, 75, blk), got)
    same = blk.copy()
    same[:, 6:8] = same[:, 4:6]
    res = ms.dmvr_refine(74, 74, same)
    assert (res[:, :2] == 0).all() and (res[:, 3] == 0).all()


def test_dmvr_errors(ms):
    import vtm_b200
    z = np.zeros((64, 64), np.int16)
    ms.upload_picture(76, z)
    ms.upload_picture(77, np.zeros((64, 48), np.int16))
    ok = np.array([[8, 8, 16, 8, 5, -3, -5, 3]], np.int32)
    assert ms.dmvr_refine(76, 76, ok).tolist() == [[0, 0, 0, 0]]
    with pytest.raises(vtm_b200.VtmmeError, match="NOPIC"):
        ms.dmvr_refine(76, 999, ok)
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):
        ms.dmvr_refine(76, 77, ok)                                   # pictures of different size
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):
        ms.dmvr_refine(76, 76, np.array([[8, 8, 4, 8, 0, 0, 0, 0]], np.int32))      # 4-wide blocks have no DMVR
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):
        ms.dmvr_refine(76, 76, np.array([[56, 8, 16, 8, 0, 0, 0, 0]], np.int32))    # outside the picture
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):
        ms.dmvr_refine(76, 76, ok, 12)


@pytest.mark.parametrize("bd", [10, 8])
def test_dmvr_final_prediction(ms, oracle_lib, bd):
    """The padded prediction after DMVR (xPrefetch + xPad + xFinalPaddedMCForDMVR) of both lists, luma and 4:2:0 chroma, with
    the refinements the GPU's own search found (integer moves of up to two samples, sub-sample steps), some blocks left
    unmoved, MVs the clip moves at the picture border — against the oracle (pinned on the reference's own members)."""
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(1200 + bd)
    W, H = 256, 192
    r0, r1, _ = make_pair(170, W, H, max_global=3, max_local=3, n_rects=4, sigma=4.0, bit_depth=bd)
    p0, p1 = pad_plane(r0), pad_plane(r1)
    c0, c1 = (np.ascontiguousarray(rng.integers(0, 1 << bd, (H // 2, W // 2), dtype=np.int16)) for _ in range(2))
    cp0, cp1 = pad_plane(c0, MARGIN // 2), pad_plane(c1, MARGIN // 2)
    ms.upload_picture(80, p0, MARGIN)
    ms.upload_picture(81, p1, MARGIN)
    ms.upload_picture(82, cp0, MARGIN // 2)
    ms.upload_picture(83, c1)                                   # chroma border of list 1 replicated on the device
    blk = dmvr_cases(rng, W, H, 500)
    mvd = np.ascontiguousarray(ms.dmvr_refine(80, 81, blk, bd)[:, :2])
    mvd[::5] = 0
    assert (np.abs(mvd) >= 16).any(axis=1).sum() > 50 and (mvd % 16 != 0).any(axis=1).sum() > 50
    stride, off = p0.shape[1], MARGIN * p0.shape[1] + MARGIN
    cstride, coff = cp0.shape[1], (MARGIN // 2) * cp0.shape[1] + MARGIN // 2
    for lst, (pid, cid, plane, cplane, sgn) in enumerate([(80, 82, p0, cp0, 1), (81, 83, p1, cp1, -1)]):
        jobs = blk.copy()
        jobs[:, 4:6] = blk[:, 4 + 2 * lst:6 + 2 * lst]                       # the list's merge MV
        jobs[:, 6:8] = jobs[:, 4:6] + sgn * mvd                              # its refined MV
        got_y = ms.dmvr_final_mc(0, pid, jobs, bd)
        got_c = ms.dmvr_final_mc(1, cid, jobs, bd)
        py = pc = 0
        for b in jobs:
            x, y, w, h, mx, my, fx, fy = (int(v) for v in b)
            want = np.zeros(w * h, np.int16)
            oracle_lib.vo_dmvr_final_luma(B.ptr(plane, off), stride, x, y, w, h, mx, my, fx, fy, W, H, 128, 128, bd, B.ptr(want))
            assert np.array_equal(got_y[py:py + w * h], want), (lst, b.tolist())
            py += w * h
            sz = (w // 2) * (h // 2)
            wantc = np.zeros(sz, np.int16)
            oracle_lib.vo_dmvr_final_chroma(B.ptr(cplane, coff), cstride, x, y, w, h, mx, my, fx, fy, W, H, 128, 128, bd, B.ptr(wantc))
            assert np.array_equal(got_c[pc:pc + sz], wantc), (lst, b.tolist())
            pc += sz
