#!/usr/bin/env python
"""DMVR search of every 16x16 sub-block of a 1080p picture: vtmme_dmvr_refine (GPU, through the C ABI, results back on the
host) against the reference's own InterPrediction members on one host core (oracle/_ref/libvtmref.so: xPrefetch, xinitMC,
xDMVRCost, xBIPMVRefine, xDMVRSubPixelErrorSurface per sub-block), with exact equality of every result.

  python microbench_dmvr.py [--reps 10] [--out profiles/r02_dmvr.md]

Workload: the two sides of a synthetic 1080p 10-bit pair (bench.py's content) as the reference pictures of list 0 / 1,
8,040 sub-blocks with random merge MVs (mirrored pair plus a mismatch of up to 1.5 samples).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
W, H = 1920, 1080


def main():
    import vtm_b200
    from oracle import bindings as B
    from vtm_b200.synth import make_pair
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    ms = vtm_b200.MotionSearch(0)
    r0, r1, _ = make_pair(0, W, H, max_global=3, max_local=4)
    ms.upload_picture(1, np.ascontiguousarray(r0))
    ms.upload_picture(2, np.ascontiguousarray(r1))
    rng = np.random.default_rng(11)
    xs, ys = np.meshgrid(np.arange(0, W - 15, 16), np.arange(0, H - 15, 16))
    n = xs.size
    mv = rng.integers(-6 * 16, 6 * 16 + 1, (n, 2))
    blk = np.ascontiguousarray(np.stack([xs.ravel(), ys.ravel(), np.full(n, 16), np.full(n, 16), mv[:, 0], mv[:, 1],
                                         -mv[:, 0] + rng.integers(-24, 25, n), -mv[:, 1] + rng.integers(-24, 25, n)],
                                        axis=1).astype(np.int32))
    got = ms.dmvr_refine(1, 2, blk)
    t0 = time.perf_counter()
    for _ in range(a.reps):
        got = ms.dmvr_refine(1, 2, blk)
    gpu_s = (time.perf_counter() - t0) / a.reps
    R = B.ref()
    m = 192
    p0, p1 = (np.ascontiguousarray(np.pad(p, m, mode="edge")) for p in (r0, r1))
    want = np.zeros((n, 4), np.int32)
    t0 = time.perf_counter()
    rc = R.ref_dmvr_blocks(B.ptr(p0), B.ptr(p1), p0.shape[1], W, H, m, n, C.c_void_p(blk.ctypes.data), 10, C.c_void_p(want.ctypes.data))
    cpu_s = time.perf_counter() - t0          # includes building the two Picture objects (a few ms)
    equal = rc == 0 and bool(np.array_equal(got, want))
    out = {"what": "DMVR search, every 16x16 sub-block of a 1080p 10-bit picture (%d sub-blocks per call)" % n,
           "gpu_seconds_per_call": gpu_s, "gpu_subblocks_per_s": n / gpu_s,
           "cpu_reference": {"kind": "reference (InterPrediction members, oracle/_ref)", "cores": 1, "seconds_per_picture": cpu_s,
                             "subblocks_per_s": n / cpu_s},
           "speedup_vs_one_core": cpu_s / gpu_s, "equal_to_reference": equal,
           "moved": int((want[:, :2] != 0).any(axis=1).sum()), "sub_sample": int((want[:, :2] % 16 != 0).any(axis=1).sum())}
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            f.write("# DMVR search micro-benchmark\n\n```json\n%s\n```\n" % json.dumps(out, indent=1))
    ms.close()
    if not equal:
        sys.exit(1)


if __name__ == "__main__":
    main()
