#!/usr/bin/env python
"""Standalone batched motion-estimation benchmark (BASELINE.json config 4).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libvtmme.so)
  python bench.py --impl reference --gpus N --steps K ...  # VTM's own CPU path (oracle/_ref) on the host cores

Workload: synthetic 1080p 10-bit frame pairs (vtm_b200/synth.py), every grid-aligned square CU 8..128
(43,020 per pair), full search SR=64 (16,641 SAD candidates per CU before border clipping) + half/quarter-pel
refinement with SATD (18 candidates).  A step = one batched search over `pairs_per_step` pairs per GPU; the pairs
of a step are independent units, ranks shard them with no collective (weak scaling: per-GPU work fixed).

metric  = ME block-candidates/s (one CU at one displacement: SAD for integer, interpolation+SATD for fractional)
value   = whole-job throughput, inputs resident in HBM;  e2e = through the C ABI with HOST buffers (H2D of both
          planes of every pair and D2H of all results inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SR, CTU, LAMBDA = 1920, 1080, 64, 128, 31.33
# dram__bytes_read.sum + dram__bytes_write.sum of me_tree_sad_kernel from one `ncu --set full` capture of a 4-pair launch
# (profiles/r01i_tree_sad_ncu.md: 40.1 MB + 489.8 MB), per pair
NCU_DRAM_BYTES_PER_PAIR = (40.096256e6 + 489.766656e6) / 4
METRIC = "ME block-candidates/s (SAD+SATD), 1080p SR=64 full search + quarter-pel"
UNIT = "block-candidates/s"


# ---- workload accounting (host, exact) ---------------------------------------------------------------------------
def _div_pow2(v, s):
    return (v + (1 << (s - 1)) - (v >= 0).astype(np.int64)) >> s


def windows(level, pred_q=None):
    """xSetSearchRange (InterSearch.cpp:3496-3535) for every CU of a level -> l, r, t, b arrays."""
    s = 8 << level
    nx, ny = WIDTH // s, HEIGHT // s
    x = (np.arange(nx, dtype=np.int64) * s)[None, :].repeat(ny, 0).ravel()
    y = (np.arange(ny, dtype=np.int64) * s)[:, None].repeat(nx, 1).ravel()
    px = np.zeros_like(x) if pred_q is None else pred_q[:, 0].astype(np.int64) * 4
    py = np.zeros_like(y) if pred_q is None else pred_q[:, 1].astype(np.int64) * 4
    hmax, hmin = (WIDTH + 8 - x - 1) * 16, (-CTU - 8 - x + 1) * 16
    vmax, vmin = (HEIGHT + 8 - y - 1) * 16, (-CTU - 8 - y + 1) * 16
    px, py = np.clip(px, hmin, hmax), np.clip(py, vmin, vmax)
    l = _div_pow2(np.clip(px - SR * 16, hmin, hmax), 4)
    r = _div_pow2(np.clip(px + SR * 16, hmin, hmax), 4)
    t = _div_pow2(np.clip(py - SR * 16, vmin, vmax), 4)
    b = _div_pow2(np.clip(py + SR * 16, vmin, vmax), 4)
    return l, r, t, b


def level_slices():
    off, acc = [], 0
    for level in range(5):
        s = 8 << level
        off.append(acc)
        acc += (WIDTH // s) * (HEIGHT // s)
    return off + [acc]


def workload_counts(pred_q=None):
    """Per pair: naive block-candidates (what the CPU evaluates) and reuse-minimal integer ops.
    pred_q: optional int16 [nCU, 2] quarter-pel predictors (run B)."""
    cands, ops, ncu = 0, 0, 0
    off = level_slices()
    for level in range(5):
        l, r, t, b = windows(level, None if pred_q is None else pred_q[off[level]:off[level + 1]])
        area = (r - l + 1) * (b - t + 1)
        n = len(area)
        ncu += n
        cands += int(area.sum()) + 18 * n
        if level == 0:
            ops += 2 * 64 * int(area.sum())      # |a-b| and accumulate per pixel-candidate at the 8x8 level
        else:
            ops += int(area.sum())               # one add per parent block-candidate
    return cands, ops, ncu


# ---- clocks -------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        mhz, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [v.strip() for v in line.split(",")]
            if len(f) < 7 or not (t0 <= ts <= t1 + 0.2):
                continue
            try:
                mhz.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


# ---- reference arm / cpu baseline: VTM's own xPatternSearch + fractional refinement on the host cores -------------------
def cpu_reference_rate(cur, ref, every, threads, steps=1, warmup=0):
    """Runs oracle/_ref (unmodified VTM behind a C shim) on every `every`-th CU of each level of one pair.
    -> (block-candidates/s, seconds per step, description, kind, threads)"""
    import ctypes as C
    from oracle import bindings as B
    lib = B.ref()
    kind = "reference"
    if lib is None:           # the compiled reference did not travel: fall back to the oracle port (single thread)
        lib, kind, threads = B.oracle(), "port", 1
    margin = 192
    refp = np.ascontiguousarray(np.pad(ref, margin, mode="edge"))
    stride = refp.shape[1]
    jobs, cands = [], 0
    for level in range(5):
        s = 8 << level
        nx = WIDTH // s
        l, r, t, b = windows(level)
        for i in range(0, len(l), every):
            x, y = (i % nx) * s, (i // nx) * s
            jobs.append(B.make_job(cur, refp, stride, (margin + y) * stride + margin + x, s, s,
                                   (int(l[i]), int(r[i]), int(t[i]), int(b[i])), (0, 0), 0, 0, 10, 1, 0, 1, LAMBDA,
                                   org_off=y * WIDTH + x, org_stride=WIDTH))
            cands += int((r[i] - l[i] + 1) * (b[i] - t[i] + 1)) + 18
    n = len(jobs)
    arr = (B.Job * n)(*jobs)
    res = (B.Result * n)()
    times = []
    for it in range(warmup + steps):
        if kind == "reference":
            dt = lib.ref_search_batch(arr, res, n, threads)
        else:
            dt = lib.vo_search_batch(arr, res, n, 0)
        if it >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    desc = "pair 0, every %d-th CU of each level (%d CUs, %.3g block-candidates) per step" % (every, n, cands)
    return cands / sec, sec, desc, kind, threads


def host_pair0():
    from vtm_b200.synth import make_pair
    cur, ref, _ = make_pair(0, WIDTH, HEIGHT)
    return cur, ref


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cur, ref = host_pair0()
    rate, sec, desc, kind, threads = cpu_reference_rate(cur, ref, args.cpu_every, threads, args.steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16/int32", "data": "synthetic",
            "config": {"workload": "config4: 1080p pairs, CUs 8..128, full search SR=64 + quarter-pel (bounded sample)",
                       "search_range": SR, "sample": desc},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---- this repo's arm ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import vtm_b200
    from vtm_b200 import FrameParams
    from vtm_b200.me import CU_RESULT_DTYPE
    from vtm_b200.peaks import int_peak
    from vtm_b200.shard import max_over_ranks, pairs_for_rank
    from vtm_b200.synth import make_pairs_torch

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.pairs_per_step, args.steps, args.warmup
    pool = max(args.pool, B)
    cands_pair, ops_pair, ncu = workload_counts()

    ms = vtm_b200.MotionSearch(local_rank)
    # a dedicated stream shared by torch (events, synthetic-data kernels) and the library (NULL would mean the
    # library's own stream, and torch's default stream handle IS NULL)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ms.set_stream(stream.cuda_stream)
    ms.set_frame_size(WIDTH, HEIGHT)

    # synthetic pairs of this rank: distinct seeds per rank, generated on the GPU, uploaded as library pictures
    ids = pairs_for_rank(world * pool, world, rank)   # pair p of the job belongs to rank p % world
    host_cur, host_ref = [], []
    for c0 in range(0, pool, 8):
        cur, ref = make_pairs_torch(ids[c0:c0 + 8], dev, WIDTH, HEIGHT)
        for i in range(cur.shape[0]):
            p = c0 + i
            ms.upload_picture_device(2 * p, cur[i].data_ptr(), WIDTH, WIDTH, HEIGHT)
            ms.upload_picture_device(2 * p + 1, ref[i].data_ptr(), WIDTH, WIDTH, HEIGHT)
            if p < args.e2e_pool:
                host_cur.append(cur[i].cpu().pin_memory())
                host_ref.append(ref[i].cpu().pin_memory())
        ms.synchronize()
        del cur, ref
    prm = FrameParams(searchRange=SR, bitDepth=10, ctuSize=CTU, lambdaMotion=LAMBDA, predSpread=0)
    d_res = torch.zeros(B * ncu * CU_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_pred_ptr, h_pred = 0, None
    if args.run == "B":
        # run B of config 4: seeded random quarter-pel predictors within +-16 px, one set per pair slot of a step
        from vtm_b200.synth import random_predictors
        h_pred = np.stack([random_predictors(7000 + i, ncu, 16) for i in range(B)])
        d_pred = torch.from_numpy(h_pred).to(dev)
        d_pred_ptr = d_pred.data_ptr()
        prm.predSpread = 33
        cands_pair = int(np.mean([workload_counts(h_pred[i])[0] for i in range(min(B, 4))]))
        ops_pair = int(np.mean([workload_counts(h_pred[i])[1] for i in range(min(B, 4))]))

    def step_ids(s):
        sel = [(s * B + i) % pool for i in range(B)]
        return [2 * p for p in sel], [2 * p + 1 for p in sel]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # INT/ALU-pipe peak of this GPU, measured now (roofline denominator; SURVEY §8d: fused |a-b|+c counts 2 ops)
    pk = int_peak(0, 1 << 16)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak_ops = 2.0 * pk["lane_instr_per_clk_per_sm"] * sms * pk["sm_mhz"] * 1e6

    for s in range(W):
        c, r = step_ids(s)
        ms.search_frames_device(c, r, prm, d_pred_ptr, d_res.data_ptr())
    ms.set_profiling(True)
    sampler = ClockSampler(local_rank)
    launches0 = ms.launches
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record(stream)
    kms = []
    for s in range(W, W + K):
        c, r = step_ids(s)
        ms.search_frames_device(c, r, prm, d_pred_ptr, d_res.data_ptr())
        kms.append(ms.frame_kernel_ms())
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    launches = ms.launches - launches0
    ms.set_profiling(False)
    elapsed_ms = ev0.elapsed_time(ev1)
    elapsed_ms = max_over_ranks(elapsed_ms, dev)
    value = world * B * K * cands_pair / (elapsed_ms * 1e-3)

    # ---- e2e: host buffers through the C ABI.  Every step uploads both planes of its pairs from page-locked host
    # memory (vtmme_upload_picture_async), searches them (vtmme_search_frames_device, asynchronous) and copies all
    # results back to page-locked host memory.  The three stages of consecutive steps overlap: pictures are
    # triple-buffered, results double-buffered; the timed region ends when the last result is on the host.
    e2e_steps = max(1, min(K, args.e2e_steps))
    npool = len(host_cur)
    res_bytes = B * ncu * CU_RESULT_DTYPE.itemsize
    d_res2 = [torch.zeros(res_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    h_res2 = [torch.empty(res_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    searched = [torch.cuda.Event() for _ in range(3)]
    copied = [torch.cuda.Event() for _ in range(2)]
    h_pred_pinned, d_pred_e2e = None, None
    if h_pred is not None:
        h_pred_pinned = torch.from_numpy(h_pred).pin_memory()
        d_pred_e2e = torch.zeros_like(h_pred_pinned, device=dev)

    def e2e_upload(s):
        """queue the H2D copies of step s on the library's copy stream; picture ids triple-buffered"""
        if s >= 3:
            searched[s % 3].synchronize()      # the search that last used this slot (step s-3) is done
        base = 100000 + (s % 3) * 2 * B
        for i in range(B):
            p = (s * B + i) % npool
            ms.upload_picture_async(base + 2 * i, host_cur[p].data_ptr(), WIDTH, WIDTH, HEIGHT)
            ms.upload_picture_async(base + 2 * i + 1, host_ref[p].data_ptr(), WIDTH, WIDTH, HEIGHT)

    def e2e_search(s):
        k = s & 1
        base = 100000 + (s % 3) * 2 * B
        if s >= 2:
            copied[k].synchronize()            # the results of step s-2 have left this buffer
        pp = 0
        if h_pred_pinned is not None:
            d_pred_e2e.copy_(h_pred_pinned, non_blocking=True)
            pp = d_pred_e2e.data_ptr()
        ms.search_frames_device([base + 2 * i for i in range(B)], [base + 2 * i + 1 for i in range(B)], prm, pp,
                                d_res2[k].data_ptr())
        searched[s % 3].record(stream)
        copy_stream.wait_event(searched[s % 3])
        with torch.cuda.stream(copy_stream):
            h_res2[k].copy_(d_res2[k], non_blocking=True)
            copied[k].record(copy_stream)

    # warm-up: every picture slot and result buffer is used once (device allocations happen here)
    e2e_upload(0)
    e2e_search(0)
    e2e_upload(1)
    e2e_search(1)
    e2e_upload(2)
    barrier()
    copy_stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    e2e_prof = os.environ.get("BENCH_E2E_PROFILE") == "1"   # development: kernel times of the last pipelined step
    if e2e_prof:
        ms.set_profiling(True)
    for s in range(2, e2e_steps + 2):
        e2e_upload(s + 1)
        e2e_search(s)
    copy_stream.synchronize()                  # the last results are on the host
    ms.synchronize()
    stream.wait_stream(copy_stream)
    e1.record(stream)
    barrier()
    if e2e_prof:
        sys.stderr.write("e2e kernel ms of the last step (tree, upper, frac): %s; step %.2f ms\n"
                         % (ms.frame_kernel_ms(), e0.elapsed_time(e1) / e2e_steps))
        ms.set_profiling(False)
    last = e2e_steps + 1
    h_res = h_res2[last & 1].numpy().view(CU_RESULT_DTYPE).reshape(B, ncu)
    # untimed: the pipelined path delivered what the synchronous host call gives for the same pictures
    base = 100000 + (last % 3) * 2 * B
    chk = ms.search_frames([base + 2 * i for i in range(B)], [base + 2 * i + 1 for i in range(B)], prm, h_pred)
    if not np.array_equal(h_res, chk):
        raise RuntimeError("e2e: pipelined results differ from the synchronous call")
    e2e_ms = max_over_ranks(e0.elapsed_time(e1), dev)
    e2e_value = world * B * e2e_steps * cands_pair / (e2e_ms * 1e-3)
    h2d = B * 2 * WIDTH * HEIGHT * 2
    d2h = B * ncu * CU_RESULT_DTYPE.itemsize

    if rank == 0:
        kms = np.array(kms)                       # [K, 3] per-step kernel durations on this rank
        k1_ms = float(kms[:, 0].mean())
        achieved = B * ops_pair / (k1_ms * 1e-3)
        # algorithmic HBM bytes of the dominant kernel per launch: both planes once + the 32x32 SAD surfaces out
        surf_bytes = B * (WIDTH // 32) * (HEIGHT // 32) * (2 * SR + 1) * (2 * SR + 8) * 4
        hbm_bytes = B * 2 * WIDTH * HEIGHT * 2 + surf_bytes
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16 samples, int32 SAD/SATD arithmetic", "data": "synthetic",
            "config": {"workload": "config4: synthetic 1080p 10-bit pairs, CUs 8..128 (43,020/pair), full search SR=64 "
                                   "+ half/quarter-pel SATD refinement, " + ("zero predictors (run A)" if args.run == "A" else "random quarter-pel predictors within +-16 px (run B)"),
                       "pairs_per_step_per_gpu": B, "pool_pairs_per_gpu": pool, "search_range": SR,
                       "block_candidates_per_pair": cands_pair, "frame_pairs_per_s": world * B * K / (elapsed_ms * 1e-3),
                       "l2": "inputs of a step (%d MB of planes + %d MB of SAD surfaces) exceed the 126 MB L2; steps cycle "
                             "through a pool of distinct pairs" % (B * 2 * 2304 * 1464 * 2 >> 20, surf_bytes >> 20)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "int_alu", "kernel": "me_tree_sad_kernel", "achieved": achieved / 1e12,
                         "peak": peak_ops / 1e12, "unit": "Tiop/s", "frac": achieved / peak_ops, "traffic": B * NCU_DRAM_BYTES_PER_PAIR,
                         "traffic_note": "DRAM bytes per launch scaled from the ncu capture in profiles/r01i_tree_sad_ncu.md; "
                                         "algorithmic bytes per launch: %d" % hbm_bytes,
                         "peak_source": "measured in this run: VABSDIFF.U32 issue rate %.1f lanes/clk/SM x %d SMs x %.0f MHz, "
                                        "fused |a-b|+c = 2 ops" % (pk["lane_instr_per_clk_per_sm"], sms, pk["sm_mhz"]),
                         "algorithmic_ops_per_launch": B * ops_pair, "kernel_ms": k1_ms,
                         "kernel_share_of_step": float(kms[:, 0].sum() / elapsed_ms),
                         "other_kernels_ms": {"me_tree_upper": float(kms[:, 1].mean()), "me_frac_frame": float(kms[:, 2].mean())},
                         "hbm": {"achieved_gbs": hbm_bytes / (k1_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "frac": hbm_bytes / (k1_ms * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s"}},
        }
        if world == 1 and not args.no_cpu:
            cur0, ref0 = host_pair0()
            rate, sec, desc, kind, threads = cpu_reference_rate(cur0, ref0, args.cpu_every, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc,
                                    "seconds": sec}
            if kind == "reference":   # the single-thread number next to the all-cores one (sparser sample)
                r1, s1, d1, _, _ = cpu_reference_rate(cur0, ref0, 8 * args.cpu_every, 1)
                line["cpu_baseline"]["single_core"] = {"value": r1, "unit": UNIT, "cores": 1, "sample": d1, "seconds": s1}
        emit(line)
    ms.close()
    if world > 1:
        dist.destroy_process_group()


_STDOUT = {"fd": None}


def emit(line):
    """The one JSON line of the bench contract, on the process's original stdout."""
    text = json.dumps(line) + "\n"
    sys.stdout.flush()
    if _STDOUT["fd"] is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_STDOUT["fd"], text.encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-step", type=int, default=32)
    ap.add_argument("--pool", type=int, default=64, help="distinct resident pairs per GPU")
    ap.add_argument("--e2e-pool", type=int, default=16, help="pairs kept in pinned host memory for the e2e pass")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--cpu-every", type=int, default=1, help="CPU baseline sample: every n-th CU of pair 0")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--run", default="A", choices=["A", "B"], help="config 4 run A (zero predictors) or B (random predictors)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything a library prints on fd 1 while we run (e.g. NCCL's version banner)
    # is sent to stderr, and emit() writes the line to the real stdout.
    sys.stdout.flush()
    _STDOUT["fd"] = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
