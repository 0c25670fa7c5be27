"""Generates tests/golden/vtm_golden.npz from the UNMODIFIED reference compiled here
(oracle/_ref/libvtmref.so, see oracle/Makefile.ref).  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures pin the oracle (tests/test_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py) on
machines where /root/reference and oracle/_ref do not exist.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

SIZES = [4, 8, 16, 32, 64, 128]


def main():
    R = B.ref()
    assert R is not None, "build the reference first: make -f oracle/Makefile.ref -j8 all"
    rng = np.random.default_rng(20261018)
    out = {"simd_level": np.array([R.ref_simd_level()])}

    # ---- distortion: every shape, uniform + residual-like, SAD (subShiftMode 0 and 2) and SATD
    d_meta, d_org, d_cur, d_val = [], [], [], []
    for w in SIZES:
        for h in SIZES:
            for kind in range(2):
                cur = rng.integers(0, 1024, (h, w), dtype=np.int16)
                if kind == 0:
                    org = rng.integers(0, 1024, (h, w), dtype=np.int16)
                else:   # bi-pred style: 2*org - otherPred
                    org = (2 * np.clip(cur + np.rint(rng.normal(0, 16, (h, w))), 0, 1023) - rng.integers(0, 1024, (h, w))).astype(np.int16)
                vals = [R.ref_dist(B.ptr(org), w, B.ptr(cur), w, w, h, 10, 0, 0),
                        R.ref_dist(B.ptr(org), w, B.ptr(cur), w, w, h, 10, 2, 0),
                        R.ref_dist(B.ptr(org), w, B.ptr(cur), w, w, h, 10, 0, 1)]
                d_meta.append((w, h, kind))
                d_org.append(org.ravel())
                d_cur.append(cur.ravel())
                d_val.append(vals)
    out["dist_meta"] = np.array(d_meta, np.int32)
    out["dist_org"] = np.concatenate(d_org)
    out["dist_cur"] = np.concatenate(d_cur)
    out["dist_val"] = np.array(d_val, np.uint64)

    # ---- MV rate
    mv = []
    for _ in range(400):
        x, y = (int(v) for v in rng.integers(-300, 300, 2))
        px, py = (int(v) for v in rng.integers(-1200, 1200, 2))
        scale, imv = int(rng.integers(0, 3)), int(rng.choice([0, 1, 2, 4]))
        mv.append((x, y, px, py, scale, imv, R.ref_mv_bits(x, y, px, py, scale, imv)))
    out["mv_bits"] = np.array(mv, np.int32)
    out["mv_cost"] = np.array([[R.ref_mv_cost(lam, b) for b in range(96)] for lam in (31.33, 4.7, 57.908)], np.uint64)
    out["mv_cost_lambda"] = np.array([31.33, 4.7, 57.908])

    # ---- interpolation: source plane 40x40, block 16x12 (and the 4x4 / 4x11 coefficient quirk), all luma phases
    src = rng.integers(0, 1024, (40, 40), dtype=np.int16)
    mid = rng.integers(-8192, 8192, (40, 40), dtype=np.int16)
    out["if_src"], out["if_mid"] = src, mid
    if_meta, if_out = [], []
    for (w, h) in [(16, 12), (4, 4), (4, 11), (9, 20)]:
        for frac in range(16):
            for (vert, first, last, alt) in [(0, 1, 0, 0), (0, 1, 1, 0), (1, 1, 1, 0), (1, 0, 1, 0), (1, 0, 0, 0), (0, 1, 0, 1), (1, 0, 1, 1)]:
                if alt and frac != 8:
                    continue
                s = src if first else mid
                dst = np.zeros((h, w), np.int16)
                if vert:
                    R.ref_filter_ver(0, B.ptr(s, 8 * 40 + 8), 40, B.ptr(dst), w, w, h, frac, first, last, 10, alt)
                else:
                    R.ref_filter_hor(0, B.ptr(s, 8 * 40 + 8), 40, B.ptr(dst), w, w, h, frac, last, 10, alt)
                if_meta.append((0, w, h, frac, vert, first, last, alt))
                if_out.append(dst.ravel())
    for frac in range(32):
        for (vert, first, last) in [(0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 0, 1)]:
            s = src if first else mid
            dst = np.zeros((8, 8), np.int16)
            if vert:
                R.ref_filter_ver(1, B.ptr(s, 8 * 40 + 8), 40, B.ptr(dst), 8, 8, 8, frac, first, last, 10, 0)
            else:
                R.ref_filter_hor(1, B.ptr(s, 8 * 40 + 8), 40, B.ptr(dst), 8, 8, 8, frac, last, 10, 0)
            if_meta.append((1, 8, 8, frac, vert, first, last, 0))
            if_out.append(dst.ravel())
    out["if_meta"] = np.array(if_meta, np.int32)
    out["if_out"] = np.concatenate(if_out)

    # ---- searches: xPatternSearch + xPatternSearchFracDIF on a 160x160 smooth plane
    H = W = 160
    base = rng.integers(0, 1024, (H + 2, W + 2)).astype(np.int32)
    plane = sum(base[dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3)) // 9
    plane = np.ascontiguousarray(plane.astype(np.int16))
    out["search_ref"] = plane
    s_meta, s_org, s_res = [], [], []
    for (w, h) in [(8, 8), (16, 16), (8, 4), (4, 8), (16, 8), (8, 16), (32, 32), (32, 8), (8, 32), (64, 64), (64, 16), (16, 64), (4, 16), (16, 4)]:
        for v in range(3):
            x, y = 48, 48
            dx, dy = (int(t) for t in rng.integers(-9, 10, 2))
            org = np.clip(plane[y + dy:y + dy + h, x + dx:x + dx + w].astype(np.int32) + np.rint(rng.normal(0, 4, (h, w))).astype(np.int32), 0, 1023).astype(np.int16)
            org = np.ascontiguousarray(org)
            pq = (int(rng.integers(-30, 31)), int(rng.integers(-30, 31)))
            win = (-12 + int(rng.integers(0, 3)), 12 - int(rng.integers(0, 3)), -12 + int(rng.integers(0, 3)), 12 - int(rng.integers(0, 3)))
            imv, alt, ssm = [(0, 0, 0), (1, 1, 0), (0, 0, 2)][v]
            lam = [31.33, 12.25, 57.9][v]
            j = B.make_job(org, plane, W, y * W + x, w, h, win, pq, imv, ssm, 10, 1, alt, 1, lam)
            r = B.Result()
            R.ref_search(C.byref(j), C.byref(r))
            s_meta.append((w, h, x, y) + win + pq + (imv, alt, ssm))
            s_org.append(org.ravel())
            s_res.append(r.tuple())
    out["search_meta"] = np.array(s_meta, np.int32)
    out["search_lambda"] = np.array([[31.33, 12.25, 57.9][i % 3] for i in range(len(s_meta))])
    out["search_org"] = np.concatenate(s_org)
    out["search_res"] = np.array(s_res, np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vtm_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
