"""Generates tests/golden/tz_selective_golden.npz from the UNMODIFIED reference compiled here (oracle/_ref/libvtmref.so):
InterSearch::xTZSearchSelective (FastSearch=2) and xTZSearch under subShiftMode 1 (xTZSearchHelp's staged SAD), each
followed by the fractional refinement at its result, called through oracle/ref_harness.cpp.  The pictures are those of
amvr_tz_golden.npz.  Run in the build container only:

    python tests/golden/make_golden_tz_selective.py

The fixture pins the oracle (tests/test_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py) where
/root/reference does not exist.  Row layout = amvr_tz_golden's tz_int plus a last column `selective`.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from tests.helpers import MARGIN, pad_plane, tz_case  # noqa: E402

W, H = 192, 128
SHAPES = [(8, 8), (16, 16), (32, 32), (64, 64), (128, 128), (16, 8), (8, 32), (64, 16), (4, 8), (32, 4)]


def main():
    R = B.ref()
    assert R is not None, "build the reference first: make -f oracle/Makefile.ref -j8 all"
    here = os.path.dirname(os.path.abspath(__file__))
    ga = np.load(os.path.join(here, "amvr_tz_golden.npz"))
    cur, ref = np.ascontiguousarray(ga["cur"]), np.ascontiguousarray(ga["ref"])
    refp = pad_plane(ref)
    stride = refp.shape[1]
    rng = np.random.default_rng(20261020)
    rows, lams, res = [], [], []
    for (w, h) in SHAPES:
        for rep in range(6):
            # reps 0-3: the selective search (staged SAD; rep 3 with RestrictMESampling, i.e. mode 2); rep 4: the cached-MV
            # fast re-search through xTZSearch with the staged SAD; rep 5: far start points at the picture corner
            selective, fast = [(1, 0), (1, 0), (1, 0), (1, 0), (0, 1), (1, 0)][rep]
            ssm = [1, 1, 1, 2, 1, 1][rep]
            x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
            y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
            if rep == 5:
                x, y = W - w, 0
            sr = [64, 32][rep & 1]
            t = tz_case(rng, x, y, W, H, sr, 0, fast, first_stop=1, max_pel=[18, 4, 40, 18, 18, 150][rep],
                        n_seeds=int(rng.integers(0, 9)))
            t.selective = selective
            pq = (int(rng.integers(-60, 61)), int(rng.integers(-60, 61)))
            lam = [31.33, 9.75][rep & 1]
            j = B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, (0, 0, 0, 0), pq, 0, ssm, 10, 1, 0, 0, lam,
                           org_off=y * W + x, org_stride=W)
            mx, my, sad = C.c_int(), C.c_int(), C.c_uint64()
            R.ref_tz_search(C.byref(j), C.byref(t), C.byref(mx), C.byref(my), C.byref(sad))
            # the reference's own xPatternSearchFracDIF body at that position (one-position xPatternSearch window)
            jf = B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, (mx.value, mx.value, my.value, my.value), pq, 0,
                            0, 10, 1, 0, 1, lam, org_off=y * W + x, org_stride=W)
            r = B.Result()
            R.ref_search(C.byref(jf), C.byref(r))
            assert (r.mvX, r.mvY) == (mx.value, my.value)
            rows.append([w, h, x, y, pq[0], pq[1], ssm, t.startX, t.startY, t.hasInt2Nx2N, t.int2Nx2NX, t.int2Nx2NY, t.nSeeds]
                        + [t.seedX[i] for i in range(16)] + [t.seedY[i] for i in range(16)]
                        + [t.searchRange, t.extended, t.fast, t.firstSearchStop, selective])
            lams.append(lam)
            res.append([mx.value, my.value, sad.value, r.halfX, r.halfY, r.qterX, r.qterY, r.fracCost])
    path = os.path.join(here, "tz_selective_golden.npz")
    np.savez_compressed(path, tz_int=np.array(rows, np.int32), tz_lambda=np.array(lams), tz_res=np.array(res, np.int64))
    print("wrote", path, os.path.getsize(path), "bytes;", len(res), "cases")


if __name__ == "__main__":
    main()
