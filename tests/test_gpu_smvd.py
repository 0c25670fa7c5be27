"""GPU: the symmetric-MVD search (vtmme_smvd_search) through the C ABI against the oracle's restatement of
InterSearch::xSymmetricMotionEstimation, which tests/test_oracle_vs_ref.py pins on the reference's own member.  Bit-exact:
both MVs and the cost of every search."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bindings as B  # noqa: E402
from tests.helpers import MARGIN, pad_plane  # noqa: E402


@pytest.fixture(scope="module")
def ms():
    import vtm_b200
    m = vtm_b200.MotionSearch(0)
    yield m
    m.close()


@pytest.mark.parametrize("imv", [0, 1, 2, 3])
def test_smvd_search(ms, oracle_lib, imv):
    """Every AMVR precision (the half-sample one with the alternative filter), SATD (8x8, 16x8 and 8x16 tilings) and SAD,
    clipped bi-prediction targets, all five BCW weights, PU shapes 8x8 .. 128x128, positions at the picture border with MVs
    the clip moves, start costs that let the diamond run several rounds or stop at once; patterns read from the uploaded
    original picture and handed over by the caller."""
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(1000 + imv)
    pic_w, pic_h = 256, 192
    n = moved = 0
    for seed in range(2):
        cur, ref0, _ = make_pair(140 + seed, pic_w, pic_h, max_global=4, max_local=6, n_rects=4, sigma=4.0)
        _, ref1, _ = make_pair(150 + seed, pic_w, pic_h, max_global=4, max_local=6, n_rects=4, sigma=4.0)
        cur = np.ascontiguousarray(cur)
        p0, p1 = pad_plane(ref0), pad_plane(ref1)
        stride, off = p0.shape[1], MARGIN * p0.shape[1] + MARGIN
        ms.upload_picture(600, cur)
        ms.upload_picture(601, p0, MARGIN)
        ms.upload_picture(602, np.ascontiguousarray(ref1))       # border replicated on the device
        ios, jobs = [], []
        for w, h in [(8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (8, 32), (64, 16), (32, 8), (128, 128), (128, 64), (16, 64)]:
            for rep in range(4):
                x = int(rng.integers(0, (pic_w - w) // 4 + 1)) * 4
                y = int(rng.integers(0, (pic_h - h) // 4 + 1)) * 4
                span = 12 * 16
                if rep == 3:
                    x, y = [0, pic_w - w][int(rng.integers(0, 2))], [0, pic_h - h][int(rng.integers(0, 2))]
                    span = 170 * 16
                io = B.SmvdIo()
                io.x, io.y, io.w, io.h, io.picW, io.picH, io.maxCuW, io.maxCuH = x, y, w, h, pic_w, pic_h, 128, 128
                io.bd, io.imv = 10, imv
                unit = [4, 16, 64, 8][imv]
                io.curPredX, io.curPredY, io.tarPredX, io.tarPredY = (int(rng.integers(-span, span + 1)) // unit * unit for _ in range(4))
                dx, dy = (int(rng.integers(-6, 7)) * unit for _ in range(2))
                io.curMvX, io.curMvY = io.curPredX + dx, io.curPredY + dy
                io.tarMvX, io.tarMvY = io.tarPredX - dx, io.tarPredY - dy
                io.clipBiPred, io.useHad = int(rep == 1), int(rep != 2)
                io.bcwIdx = [2, 2, 2, 2, 0, 1, 3, 4][len(ios) % 8]
                io.lambda_ = [31.33, 8.5, 57.9, 31.33][rep]
                io.cost = [2 ** 40, w * h * 12, w * h * 5, 2 ** 40][rep]
                ios.append(io)
                job = {"curPic": 600, "refPicCur": 601, "refPicTar": 602, "x": x, "y": y, "w": w, "h": h, "imv": imv,
                       "curPred": (io.curPredX, io.curPredY), "tarPred": (io.tarPredX, io.tarPredY), "curMv": (io.curMvX, io.curMvY),
                       "tarMv": (io.tarMvX, io.tarMvY), "clipBiPred": io.clipBiPred, "useHad": io.useHad, "bcwIdx": io.bcwIdx,
                       "lambdaMotion": io.lambda_, "cost": io.cost}
                if rep == 2:
                    job["org"] = cur[y:y + h, x:x + w]            # the caller's own pattern buffer
                jobs.append(job)
        got = ms.smvd_search(jobs)
        for i, io in enumerate(ios):
            start = io.tuple()
            oracle_lib.vo_smvd_search(B.ptr(cur, io.y * pic_w + io.x), pic_w, B.ptr(p0, off), B.ptr(p1, off), stride, C.byref(io))
            assert got[i] == io.tuple(), (seed, i, io.w, io.h, start, got[i], io.tuple())
            n += 1
            moved += io.tuple()[:2] != start[:2]
    assert n == 88 and moved > 40
