"""Generates tests/golden/amvr_tz_golden.npz from the UNMODIFIED reference compiled here (oracle/_ref/libvtmref.so):
InterSearch::xTZSearch (+ the fractional refinement at its result) and xPatternSearch + xPatternSearchIntRefine, called
through oracle/ref_harness.cpp.  Run in the build container only:

    python tests/golden/make_golden_amvr_tz.py

The fixtures pin the oracle (tests/test_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py) where
/root/reference does not exist.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from tests.helpers import MARGIN, int_refine_case, pad_plane, tz_case  # noqa: E402

W, H = 192, 128
SHAPES = [(8, 8), (16, 16), (32, 32), (64, 64), (128, 128), (16, 8), (8, 32), (64, 16), (4, 8), (32, 4)]


def main():
    from vtm_b200.synth import make_pair
    R = B.ref()
    assert R is not None, "build the reference first: make -f oracle/Makefile.ref -j8 all"
    rng = np.random.default_rng(20261019)
    cur, ref, _ = make_pair(777, W, H, max_global=14, max_local=20, n_rects=3, sigma=5.0)
    cur, ref = np.ascontiguousarray(cur), np.ascontiguousarray(ref)
    refp = pad_plane(ref)
    stride = refp.shape[1]
    out = {"cur": cur, "ref": ref}

    # ---- TZ search: FastSearch=1, enhanced, fast re-search; FEN sub-sampling; seeds; 2Nx2N MV
    tz_int, tz_lam, tz_res = [], [], []
    for (w, h) in SHAPES:
        for rep in range(6):
            extended, fast = [(0, 0), (1, 0), (0, 1)][rep % 3]
            x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
            y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
            if rep == 5:
                x, y = W - w, 0
            sr = [64, 32][rep & 1]
            t = tz_case(rng, x, y, W, H, sr, extended, fast, first_stop=int(rep != 4), max_pel=18 if rep < 5 else 150, n_seeds=int(rng.integers(0, 9)))
            pq = (int(rng.integers(-60, 61)), int(rng.integers(-60, 61)))
            ssm = 2 if rep >= 3 else 0
            lam = [31.33, 9.75][rep & 1]
            j = B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, (0, 0, 0, 0), pq, 0, ssm, 10, 1, 0, 0, lam,
                           org_off=y * W + x, org_stride=W)
            mx, my, sad = C.c_int(), C.c_int(), C.c_uint64()
            R.ref_tz_search(C.byref(j), C.byref(t), C.byref(mx), C.byref(my), C.byref(sad))
            # fractional refinement at the TZ result: a one-position xPatternSearch window makes the reference's own
            # xPatternSearchFracDIF body (ref_search) start there
            jf = B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, (mx.value, mx.value, my.value, my.value), pq, 0,
                            ssm, 10, 1, 0, 1, lam, org_off=y * W + x, org_stride=W)
            r = B.Result()
            R.ref_search(C.byref(jf), C.byref(r))
            assert (r.mvX, r.mvY) == (mx.value, my.value)
            tz_int.append([w, h, x, y, pq[0], pq[1], ssm, t.startX, t.startY, t.hasInt2Nx2N, t.int2Nx2NX, t.int2Nx2NY, t.nSeeds]
                          + [t.seedX[i] for i in range(16)] + [t.seedY[i] for i in range(16)]
                          + [t.searchRange, t.extended, t.fast, t.firstSearchStop])
            tz_lam.append(lam)
            tz_res.append([mx.value, my.value, sad.value, r.halfX, r.halfY, r.qterX, r.qterY, r.fracCost])
    out["tz_int"] = np.array(tz_int, np.int32)
    out["tz_lambda"] = np.array(tz_lam)
    out["tz_res"] = np.array(tz_res, np.int64)

    # ---- integer / 4-pel AMVR: xPatternSearch over a window, then xPatternSearchIntRefine
    ir_int, ir_f, ir_res = [], [], []
    for (w, h) in SHAPES:
        for rep in range(4):
            imv = 1 + (rep & 1)
            use_had = int(rep != 3)
            x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
            y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
            if rep == 2:
                x, y = 0, H - h
            io = int_refine_case(rng, imv, x, y, w, h, W, H, max_pel=5)
            pred16 = (io.candX[io.mvpIdx], io.candY[io.mvpIdx])
            pq = tuple((v + 1) >> 2 if v >= 0 else (v + 2) >> 2 for v in pred16)
            l, r_, t_, b = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            B.oracle().vo_set_search_range(pq[0] * 4, pq[1] * 4, x, y, W, H, 128, 128, 6, C.byref(l), C.byref(r_), C.byref(t_), C.byref(b))
            win = (l.value, r_.value, t_.value, b.value)
            lam = [31.33, 14.5][rep & 1]
            j = B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, win, pq, imv << 1, 0, 10, use_had, 0, 0, lam,
                           org_off=y * W + x, org_stride=W)
            r = B.Result()
            R.ref_search(C.byref(j), C.byref(r))
            io.mvX, io.mvY = r.mvX * 16, r.mvY * 16
            row = [w, h, x, y, pq[0], pq[1], win[0], win[1], win[2], win[3], imv, use_had, io.numCand, io.candX[0], io.candY[0],
                   io.candX[1], io.candY[1], io.mvpIdx, io.mvpIdxBits[0], io.mvpIdxBits[1], io.bits]
            fw = io.fWeight
            R.ref_int_refine(C.byref(j), C.byref(io))
            ir_int.append(row)
            ir_f.append([lam, fw])
            ir_res.append([r.mvX, r.mvY, r.intSad, io.mvX, io.mvY, io.mvpIdx, io.bits, io.cost])
    out["ir_int"] = np.array(ir_int, np.int32)
    out["ir_f"] = np.array(ir_f)
    out["ir_res"] = np.array(ir_res, np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "amvr_tz_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(tz_res), "TZ cases,", len(ir_res), "AMVR cases")


if __name__ == "__main__":
    main()
