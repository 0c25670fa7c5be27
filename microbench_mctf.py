#!/usr/bin/env python
"""Motion estimation of the GOP-based temporal filter on 1080p pictures: vtmme_mctf_me (GPU, through the C ABI, vectors
back on the host) against the reference's own EncTemporalFilter::motionEstimation on one host core
(oracle/_ref/libvtmref.so), with exact equality of every vector and error.

  python microbench_mctf.py [--refs 4] [--reps 5] [--out profiles/r01h_mctf.md]

Workload: one original 1080p 10-bit picture against `refs` neighbouring pictures (the filter uses up to 4 per side),
synthetic content of bench.py (config 4).
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
    ap.add_argument("--refs", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    ms = vtm_b200.MotionSearch(0)
    cur, ref0, _ = make_pair(0, W, H)
    refs = [ref0] + [make_pair(k, W, H, max_global=10 + 6 * k, max_local=20)[0] for k in range(1, a.refs)]
    ms.upload_picture(1, np.ascontiguousarray(cur))
    for k, r in enumerate(refs):
        ms.upload_picture(10 + k, np.ascontiguousarray(r))
    ids_o, ids_r = [1] * a.refs, [10 + k for k in range(a.refs)]
    got = ms.mctf_me(ids_o, ids_r, W, H, 10)
    t0 = time.perf_counter()
    for _ in range(a.reps):
        got = ms.mctf_me(ids_o, ids_r, W, H, 10)
    gpu_s = (time.perf_counter() - t0) / a.reps
    # the reference on one core, first reference picture; equality of the whole field
    R = B.ref()
    curp = np.ascontiguousarray(np.pad(cur, 128, mode="edge"))
    refp = np.ascontiguousarray(np.pad(refs[0], 128, mode="edge"))
    stride = curp.shape[1]
    off = 128 * stride + 128
    want = np.zeros((H // 4, W // 4, 3), np.int32)
    cpu_s = R.ref_mctf_me(B.ptr(curp, off), stride, B.ptr(refp, off), stride, W, H, 10, C.c_void_p(want.ctypes.data))
    equal = bool(np.array_equal(got[0], want))
    nblk = ((W - 1) // 8) * ((H - 1) // 8)
    out = {"what": "GOP-based temporal filter motion estimation, 1080p 10-bit, %d reference pictures per call" % a.refs,
           "gpu_seconds_per_call": gpu_s, "gpu_frame_pairs_per_s": a.refs / gpu_s, "gpu_blocks8x8_per_s": a.refs * nblk / gpu_s,
           "cpu_reference": {"kind": "reference (EncTemporalFilter::motionEstimation, oracle/_ref)", "cores": 1,
                             "seconds_per_pair": cpu_s, "frame_pairs_per_s": 1.0 / cpu_s},
           "speedup_vs_one_core": (a.refs / gpu_s) * cpu_s, "equal_to_reference": equal, "compared_blocks": nblk}
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            f.write("# Temporal-filter motion estimation micro-benchmark\n\n```json\n%s\n```\n" % json.dumps(out, indent=1))
    ms.close()
    if not equal:
        sys.exit(1)


if __name__ == "__main__":
    main()
