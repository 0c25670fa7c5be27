#!/usr/bin/env python
"""Batched TZ search (FastSearch=1 / 3) on whole 1080p pictures: the GPU frame search (vtmme_search_frames with
fastSearch, one warp per CU) against the reference's own InterSearch::xTZSearch on the host cores
(oracle/_ref/libvtmref.so), with exact equality of every integer MV and SAD of one whole picture pair.

  python microbench_tz.py [--pairs 32] [--steps 4] [--fast-search 1] [--out profiles/r01g_tz.md]

Workload: the synthetic 1080p 10-bit pairs of bench.py (config 4), every grid-aligned square CU 8..128 (43,020 per
pair), each searched from the zero predictor with SearchRange 64, then half/quarter-pel SATD refinement.
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
WIDTH, HEIGHT, SR, LAMBDA, MARGIN = 1920, 1080, 64, 31.33, 192


def cpu_tz(cur, ref, fast_search, threads, every=1):
    """The reference's xTZSearch for every `every`-th CU of one pair -> (mv [n,2], sad [n], cu indices, seconds)."""
    from oracle import bindings as B
    R = B.ref()
    refp = np.ascontiguousarray(np.pad(ref, MARGIN, mode="edge"))
    stride = refp.shape[1]
    jobs, tzs, idx = [], [], []
    cu = 0
    for level in range(5):
        s = 8 << level
        for cy in range(HEIGHT // s):
            for cx in range(WIDTH // s):
                if cu % every == 0:
                    x, y = cx * s, cy * s
                    jobs.append(B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, s, s, (0, 0, 0, 0), (0, 0),
                                           0, 0, 10, 1, 0, 0, LAMBDA, org_off=y * WIDTH + x, org_stride=WIDTH))
                    t = B.TzParams()
                    t.searchRange, t.extended, t.fast, t.firstSearchStop = SR, int(fast_search == 3), 0, 1
                    t.posX, t.posY, t.picW, t.picH, t.maxCuW, t.maxCuH = x, y, WIDTH, HEIGHT, 128, 128
                    tzs.append(t)
                    idx.append(cu)
                cu += 1
    n = len(jobs)
    ja, ta = (B.Job * n)(*jobs), (B.TzParams * n)(*tzs)
    mv = np.zeros((n, 2), np.int32)
    sad = np.zeros(n, np.uint64)
    sec = R.ref_tz_batch(ja, ta, n, threads, C.c_void_p(mv.ctypes.data), C.c_void_p(sad.ctypes.data))
    return mv, sad, np.array(idx), sec


def main():
    import torch
    import vtm_b200
    from vtm_b200 import FrameParams
    from vtm_b200.me import CU_RESULT_DTYPE
    from vtm_b200.synth import make_pair, make_pairs_torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=32)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--fast-search", type=int, default=1, choices=[1, 3])
    ap.add_argument("--cpu-every", type=int, default=1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    ms = vtm_b200.MotionSearch(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ms.set_stream(stream.cuda_stream)
    ncu = ms.set_frame_size(WIDTH, HEIGHT)
    B = a.pairs
    for c0 in range(0, 2 * B, 8):
        cur, ref = make_pairs_torch(list(range(c0, c0 + 8)), dev, WIDTH, HEIGHT)
        for i in range(cur.shape[0]):
            ms.upload_picture_device(2 * (c0 + i), cur[i].data_ptr(), WIDTH, WIDTH, HEIGHT)
            ms.upload_picture_device(2 * (c0 + i) + 1, ref[i].data_ptr(), WIDTH, WIDTH, HEIGHT)
        ms.synchronize()
    prm = FrameParams(searchRange=SR, lambdaMotion=LAMBDA, fastSearch=a.fast_search)
    d_res = torch.zeros(B * ncu * CU_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)

    def ids(s):
        sel = [(s * B + i) % (2 * B) for i in range(B)]
        return [2 * p for p in sel], [2 * p + 1 for p in sel]

    for s in range(2):
        ms.search_frames_device(*ids(s), prm, 0, d_res.data_ptr())
    ms.set_profiling(True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    kms = []
    for s in range(2, 2 + a.steps):
        ms.search_frames_device(*ids(s), prm, 0, d_res.data_ptr())
        kms.append(ms.frame_kernel_ms())
    e1.record(stream)
    torch.cuda.synchronize()
    ms.set_profiling(False)
    step_ms = e0.elapsed_time(e1) / a.steps
    kms = np.array(kms)

    # equality on one whole pair (host-generated, the reference sees the same samples)
    cur0, ref0, _ = make_pair(0, WIDTH, HEIGHT)
    ms.upload_picture(900000, cur0)
    ms.upload_picture(900001, np.ascontiguousarray(np.pad(ref0, MARGIN, mode="edge")), MARGIN)
    got = ms.search_frames([900000], [900001], prm)[0]
    threads = os.cpu_count() or 1
    mv, sad, idx, sec = cpu_tz(cur0, ref0, a.fast_search, threads, a.cpu_every)
    equal = bool(np.array_equal(got["intX"][idx], mv[:, 0]) and np.array_equal(got["intY"][idx], mv[:, 1])
                 and np.array_equal(got["intSad"][idx].astype(np.uint64), sad))
    out = {"what": "batched TZ search (FastSearch=%d), 1080p, 43,020 CUs per pair, SR=64, zero predictors" % a.fast_search,
           "pairs_per_step": B, "steps": a.steps, "ms_per_step": step_ms,
           "frame_pairs_per_s": B / (step_ms * 1e-3), "cu_searches_per_s": B * ncu / (step_ms * 1e-3),
           "kernel_ms": {"me_tz_frame": float(kms[:, 0].mean()), "me_frac_frame": float(kms[:, 2].mean())},
           "cpu_reference": {"kind": "reference (oracle/_ref xTZSearch, integer search only)", "cores": threads,
                             "searches": int(len(idx)), "seconds": sec, "cu_searches_per_s": len(idx) / sec},
           "gpu_tz_kernel_cu_searches_per_s": B * ncu / (float(kms[:, 0].mean()) * 1e-3),
           "equal_to_reference": equal, "compared_cus": int(len(idx))}
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            f.write("# Batched TZ search micro-benchmark\n\n```json\n%s\n```\n" % json.dumps(out, indent=1))
    ms.close()
    if not equal:
        sys.exit(1)


if __name__ == "__main__":
    main()
