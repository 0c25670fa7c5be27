#!/usr/bin/env python
"""Distortion / interpolation micro-benchmark (BASELINE config 5): SAD (subShift 0/1), SATD and the 8-tap luma /
4-tap chroma filters over every VVC CU size, GPU batched kernels (libvtmme table-level entry points, device-resident
blocks) against the reference's own AVX2 function pointers on one host core (oracle/_ref/libvtmref.so), with
exact equality of every output.

  python microbench.py [--n 65536] [--out profiles/microbench.md] [--quick]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
SIZES = [4, 8, 16, 32, 64, 128]


def main():
    import torch
    import vtm_b200
    from oracle import bindings as B
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time direct launches (host call overhead included)")
    a = ap.parse_args()
    R = B.ref()
    O = B.oracle()
    ms = vtm_b200.MotionSearch(0)
    stream = torch.cuda.Stream()          # not the default stream: its handle is NULL = "library stream" in the C ABI
    torch.cuda.set_stream(stream)
    ms.set_stream(stream.cuda_stream)
    rows = []
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        peak_src = "MEASURED_PEAKS.json"
    except (OSError, ValueError, KeyError):
        hbm_peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    L2_BYTES = 126e6
    shapes = [(w, h) for w in SIZES for h in SIZES]
    if a.quick:
        shapes = [(8, 8), (16, 16), (32, 8), (64, 64), (128, 128), (4, 4)]

    def gpu_time(fn, reps=5):
        """Kernel time per call: the calls are captured into a CUDA graph (a ctypes call costs ~10 us of host time, more than a
        small batch takes on the device) and the graph is replayed between two events; direct launches if capture fails."""
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        graph = None
        if not a.no_graph:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    for _ in range(reps):
                        fn()
                graph = g
            except Exception as exc:      # noqa: BLE001 - fall back to plain launches
                print("graph capture failed (%s): timing direct launches" % exc, file=sys.stderr)
                torch.cuda.synchronize()
        torch.cuda.set_stream(stream)
        if graph is not None:
            graph.replay()
            torch.cuda.synchronize()
            e0.record(stream)
            graph.replay()
            e1.record(stream)
        else:
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    for (w, h) in shapes:
        n = min(a.n, (1 << 28) // (w * h * 2))          # cap the batch at 256 MB per operand
        rng = np.random.default_rng(7 * w + h)
        for dist_name in ("uniform", "residual"):
            org = rng.integers(0, 1024, (n, h, w), dtype=np.int16)
            if dist_name == "uniform":
                cur = rng.integers(0, 1024, (n, h, w), dtype=np.int16)
            else:
                cur = np.clip(org + np.rint(rng.normal(0, 16, org.shape)), 0, 1023).astype(np.int16)
            d_org, d_cur = torch.from_numpy(org).cuda(), torch.from_numpy(cur).cuda()
            d_out = torch.zeros(n, dtype=torch.int64, device="cuda")
            for (kind, mode, label) in ((0, 0, "SAD"), (0, 2, "SAD subShift(FEN)"), (1, 0, "SATD")):
                if (w == 4 and h == 4 and kind == 0 and mode == 2):
                    continue
                ss = O.vo_subshift(mode, w, h) if kind == 0 else 0
                t_gpu = gpu_time(lambda: ms.dist_batch(kind, d_org.data_ptr(), w, w * h, d_cur.data_ptr(), w, w * h, w, h, ss, n, d_out.data_ptr()))
                got = d_out.cpu().numpy().astype(np.uint64)
                ref_out = np.zeros(n, np.uint64)
                if R is not None:
                    t_cpu = R.ref_dist_batch(B.ptr(org.reshape(-1)), B.ptr(cur.reshape(-1)), w, h, n, 10, mode, kind, C.c_void_p(ref_out.ctypes.data))
                    src = "reference AVX2, 1 core"
                else:
                    t_cpu, src = float("nan"), "n/a"
                    for i in range(0, n, 997):
                        ref_out[i] = O.vo_sad(B.ptr(org[i]), w, B.ptr(cur[i]), w, w, h, ss) if kind == 0 else O.vo_satd(B.ptr(org[i]), w, B.ptr(cur[i]), w, w, h)
                    got = got.copy()
                    mask = np.ones(n, bool)
                    mask[::997] = False
                    got[mask] = 0
                equal = bool(np.array_equal(got, ref_out))
                # algorithmic bytes: the sampled rows of both blocks in, one 8-byte distortion out
                nbytes = n * (2 * w * (h >> ss) * 2 + 8)
                rows.append({"op": label, "w": w, "h": h, "data": dist_name, "n": n, "gpu_blocks_per_s": n / t_gpu,
                             "cpu_blocks_per_s": n / t_cpu, "gpu_GBps": nbytes / t_gpu / 1e9, "hbm_frac": nbytes / t_gpu / 1e9 / hbm_peak,
                             "l2_resident": 2 * n * w * h * 2 < L2_BYTES, "equal": equal, "cpu": src})
                print(json.dumps(rows[-1]))
                assert equal, (label, w, h, dist_name)
        # interpolation: luma H / V / HV(second stage) and chroma, all phases summed into one timing
        nf = min(n, 16384)
        src = rng.integers(0, 1024, (nf, h + 8, w + 8), dtype=np.int16)
        mid = rng.integers(-8192, 8192, (nf, h + 8, w + 8), dtype=np.int16)
        d_src, d_mid = torch.from_numpy(src).cuda(), torch.from_numpy(mid).cuda()
        d_dst = torch.zeros((nf, h, w), dtype=torch.int16, device="cuda")
        off = 4 * (w + 8) + 4
        for (comp, vert, first, last, label, fracs) in ((0, 0, 1, 0, "luma8 hor (first)", range(1, 16)),
                                                       (0, 1, 0, 1, "luma8 ver (last)", range(1, 16)),
                                                       (0, 1, 1, 1, "luma8 ver (single)", range(1, 16)),
                                                       (1, 0, 1, 1, "chroma4 hor (single)", range(1, 32)),
                                                       (1, 1, 0, 1, "chroma4 ver (last)", range(1, 32))):
            if a.quick:
                fracs = list(fracs)[::5]
            s_np, s_d = (src, d_src) if first else (mid, d_mid)
            t_gpu = t_cpu = 0.0
            equal = True
            fracs = list(fracs)
            taps = 8 if comp == 0 else 4
            # algorithmic bytes per block: the block plus its tap halo along the filtered axis in, the block out
            blk_bytes = ((w + taps - 1) * h if not vert else w * (h + taps - 1)) * 2 + w * h * 2
            for frac in fracs:
                t_gpu += gpu_time(lambda: ms.interp_batch(comp, vert, s_d.data_ptr() + 2 * off, w + 8, (w + 8) * (h + 8), d_dst.data_ptr(), w, w * h, w, h, frac, first, last, 10, 0, nf), reps=2)
                got = d_dst.cpu().numpy()
                if R is not None:
                    want = np.zeros((nf, h, w), np.int16)
                    t_cpu += R.ref_filter_batch(comp, vert, B.ptr(s_np.reshape(-1)), B.ptr(want.reshape(-1)), w, h, nf, frac, first, last, 10)
                    equal &= bool(np.array_equal(got, want))
            rows.append({"op": label, "w": w, "h": h, "data": "uniform", "n": nf * len(list(fracs)), "gpu_blocks_per_s": nf * len(list(fracs)) / t_gpu,
                         "cpu_blocks_per_s": (nf * len(list(fracs)) / t_cpu) if t_cpu else float("nan"), "equal": equal,
                         "gpu_GBps": nf * len(fracs) * blk_bytes / t_gpu / 1e9, "hbm_frac": nf * len(fracs) * blk_bytes / t_gpu / 1e9 / hbm_peak,
                         "l2_resident": nf * (w + 8) * (h + 8) * 2 < L2_BYTES,
                         "cpu": "reference AVX2, 1 core" if R is not None else "n/a"})
            print(json.dumps(rows[-1]))
            assert equal, (label, w, h)
    mc_rows = bench_mc(ms, torch, stream, B, R, O, a.quick)
    if a.out:
        with open(a.out, "w") as f:
            f.write("# Distortion / interpolation micro-benchmark (BASELINE config 5)\n\n")
            f.write("GPU: libvtmme table-level batch kernels, blocks resident in HBM, CUDA-event timed%s.  CPU: the reference's own "
                    "dispatch-table entries (AVX2) on ONE host core.  Every output compared for exact equality.\n\n"
                    % ("" if a.no_graph else " (the calls of a timing loop are captured into a CUDA graph and replayed, so the ~10 us of host "
                       "time per ctypes call does not hide the kernels of small batches)"))
            f.write("Roofline: these kernels are HBM-bound by design (a few integer ops per byte).  `GB/s` = algorithmic bytes (operand "
                    "blocks in — sampled rows only for SAD, with the tap halo for filters — plus results out) / kernel time; `frac` = "
                    "that over the measured HBM peak of %.0f GB/s (%s).  Batches whose operands fit the 126 MB L2 are marked `L2`: "
                    "their repeated timing runs out of L2, so `frac` can exceed 1 there and is not an HBM figure.\n\n" % (hbm_peak, peak_src))
            f.write("| op | WxH | data | blocks | GPU blocks/s | CPU blocks/s (1 core) | ratio | GB/s | frac | equal |\n|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write("| %s | %dx%d | %s | %d | %.3g | %.3g | %.1f | %.0f | %.2f%s | %s |\n" % (
                    r["op"], r["w"], r["h"], r["data"], r["n"], r["gpu_blocks_per_s"], r["cpu_blocks_per_s"],
                    r["gpu_blocks_per_s"] / r["cpu_blocks_per_s"], r["gpu_GBps"], r["hbm_frac"], " (L2)" if r["l2_resident"] else "", r["equal"]))
            if mc_rows:
                f.write("\n## Motion compensation of a whole 1080p picture (xPredInterBlk, uni-directional, random fractional MVs)\n\n"
                        "GPU: `vtmme_mc_batch` (host block list -> device prediction, wall time per call incl. the host-side tile "
                        "build); CPU: the reference's own `InterPrediction::xPredInterBlk` on one core.\n\n"
                        "| component | block | blocks | GPU Msamples/s | CPU Msamples/s (1 core) | ratio | equal |\n|---|---|---|---|---|---|---|\n")
                for r in mc_rows:
                    f.write("| %s | %dx%d | %d | %.0f | %.0f | %.1f | %s |\n" % (r["comp"], r["s"], r["s"], r["n"], r["gpu_msps"],
                                                                                   r["cpu_msps"], r["gpu_msps"] / r["cpu_msps"] if r["cpu_msps"] else float("nan"), r["equal"]))
    ms.close()


def bench_mc(ms, torch, stream, B, R, O, quick):
    """Whole-picture motion compensation: every SxS block of a 1080p luma plane (960x540 chroma plane) with its own
    random fractional MV."""
    import time
    from vtm_b200.lib import CMcBlock
    rows = []
    W, H = 1920, 1080
    for comp, name in ((0, "luma 8-tap"), (1, "chroma 4-tap")):
        cw, ch = (W, H) if comp == 0 else (W // 2, H // 2)
        M = 192 if comp == 0 else 128      # the reference Picture's own chroma margin is 144
        rng = np.random.default_rng(90 + comp)
        padded = np.ascontiguousarray(np.pad(rng.integers(0, 1024, (ch, cw), dtype=np.int16), M, mode="edge"))
        ms.upload_picture(900 + comp, padded, M)
        for s in ((8, 16, 64) if not quick else (16,)):
            if comp == 1:
                s //= 2
            blks = [(x, y, min(s, cw - x), min(s, ch - y), int(rng.integers(-32 * 16, 32 * 16)), int(rng.integers(-32 * 16, 32 * 16)))
                    for y in range(0, ch, s) for x in range(0, cw, s)]
            arr = (CMcBlock * len(blks))()
            for i, b in enumerate(blks):
                arr[i] = CMcBlock(900 + comp, b[0], b[1], b[2], b[3], b[4], b[5], 0)
            total = sum(b[2] * b[3] for b in blks)
            d_dst = torch.zeros(total, dtype=torch.int16, device="cuda")
            for _ in range(2):
                ms.mc_batch(comp, arr, d_dst.data_ptr(), 0)
            ms.synchronize()
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                ms.mc_batch(comp, arr, d_dst.data_ptr(), 0)
            ms.synchronize()
            t_gpu = (time.perf_counter() - t0) / reps
            got = d_dst.cpu().numpy()
            if R is not None:
                ba = np.array(blks, dtype=np.int32)
                want = np.zeros(total, np.int16)
                sec = C.c_double()
                assert R.ref_mc_blocks(comp, B.ptr(padded), padded.shape[1], W, H, M, len(blks), C.c_void_p(ba.ctypes.data), 0, 10, 0,
                                       B.ptr(want), C.byref(sec)) == 0
                t_cpu = sec.value
                equal = bool(np.array_equal(got, want))
            else:
                t_cpu, equal = float("nan"), True
                stride = padded.shape[1]
                off = 0
                for i, (x, y, w, h, mvx, mvy) in enumerate(blks):
                    if i % 97 == 0:
                        d = np.zeros((h, w), np.int16)
                        O.vo_mc_block(comp, B.ptr(padded, (M + y) * stride + M + x), stride, w, h, mvx, mvy, 0, 10, 0, B.ptr(d), w)
                        equal &= bool(np.array_equal(d.ravel(), got[off:off + w * h]))
                    off += w * h
            rows.append({"op": "mc", "comp": name, "s": s, "n": len(blks), "gpu_msps": total / t_gpu / 1e6,
                         "cpu_msps": total / t_cpu / 1e6 if t_cpu == t_cpu else float("nan"), "equal": equal})
            print(json.dumps(rows[-1]))
            assert equal, ("mc", name, s)
    return rows


if __name__ == "__main__":
    main()
