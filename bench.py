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
TOTAL_PAIRS = 256          # BASELINE config 4: 256 frame pairs, pair p on rank p % world
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


def _union_area(rects):
    """Exact area of the union of k inclusive integer rectangles per row: rects int64 [n, k, 4] = (l, r, t, b); empty
    rectangles have r < l.  Sweep over the (at most 2k - 1) x-intervals between sorted edges."""
    n, k, _ = rects.shape
    l, r, t, b = rects[..., 0], rects[..., 1] + 1, rects[..., 2], rects[..., 3] + 1     # half-open
    ok = (r > l) & (b > t)
    xs = np.sort(np.concatenate([np.where(ok, l, 0), np.where(ok, r, 0)], axis=1), axis=1)   # [n, 2k]
    area = np.zeros(n, dtype=np.int64)
    for i in range(2 * k - 1):
        x0, x1 = xs[:, i], xs[:, i + 1]
        cover = ok & (l <= x0[:, None]) & (r >= x1[:, None])                                # rectangles spanning the interval
        # union of the y-intervals of the covering rectangles: sort by start, sweep
        ts = np.where(cover, t, np.iinfo(np.int64).max // 2)
        order = np.argsort(ts, axis=1)
        ts = np.take_along_axis(ts, order, axis=1)
        bs = np.take_along_axis(np.where(cover, b, np.iinfo(np.int64).min // 2), order, axis=1)
        length = np.zeros(n, dtype=np.int64)
        cur_end = np.full(n, np.iinfo(np.int64).min // 2)
        for j in range(k):
            live = bs[:, j] > ts[:, j]
            start = np.maximum(ts[:, j], cur_end)
            length += np.where(live & (bs[:, j] > start), bs[:, j] - start, 0)
            cur_end = np.where(live, np.maximum(cur_end, bs[:, j]), cur_end)
        area += length * (x1 - x0)
    return area


def tree_needed_ops(pred_q):
    """Integer ops the quad-tree reuse cannot avoid with per-CU predictors (run B): the SAD of an 8x8 block is needed at
    every displacement in the window of the block OR of any of its four ancestors (their SADs are sums over it), i.e. on
    the union of five windows instead of one; plus the one add per parent block-candidate of workload_counts()."""
    off = level_slices()
    nx0, ny0 = WIDTH // 8, HEIGHT // 8
    cx = np.tile(np.arange(nx0), ny0)
    cy = np.repeat(np.arange(ny0), nx0)
    rects = np.zeros((nx0 * ny0, 5, 4), dtype=np.int64)
    rects[:, :, 1] = -1
    adds = 0
    for level in range(5):
        s = 8 << level
        nx, ny = WIDTH // s, HEIGHT // s
        l, r, t, b = windows(level, None if pred_q is None else pred_q[off[level]:off[level + 1]])
        if level:
            adds += int(((r - l + 1) * (b - t + 1)).sum())
        ax, ay = cx >> level, cy >> level
        ok = (ax < nx) & (ay < ny)
        idx = np.where(ok, ay * nx + ax, 0)
        for k, arr in enumerate((l, r, t, b)):
            rects[:, level, k] = np.where(ok, arr[idx], rects[:, level, k])
    return 2 * 64 * int(_union_area(rects).sum()) + adds


def frac_ops_per_pair():
    """Algorithmic integer ops of the half + quarter-pel refinement of every CU of a pair (16 ops per 8-tap
    output sample, SATD n*(3+log2 n) per 8x8 tile of n = 64 samples, 9 candidates per stage).  Half-pel stage: 2 horizontal
    planes of (W+1)x(H+8), 4 vertical planes (W x H .. (W+1)x(H+1)); quarter-pel stage: 2 horizontal planes of W x (H+8),
    8 vertical planes of W x H; the frac-0 'filters' of the reference are copies and count as nothing."""
    ops = 0
    for level in range(5):
        s = 8 << level
        n = (WIDTH // s) * (HEIGHT // s)
        half = 16 * (1 * (s + 1) * (s + 8) + (s * (s + 1) + (s + 1) * (s + 1)))         # hor: frac 8; ver: the 2 frac-8 planes
        qter = 16 * (2 * s * (s + 8) + 8 * s * s)
        satd = 18 * (s // 8) * (s // 8) * 64 * (3 + 6)
        ops += n * (half + qter + satd)
    return ops


def ncu_traffic_per_pair():
    """DRAM bytes of the dominant kernel per frame pair from the committed `ncu --set full` capture (profiles/ncu_traffic.json,
    written by scripts/ncu_summary.py) -> (bytes per pair, source) or (None, reason)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["pairs_per_launch"], t["source"]
    except (OSError, ValueError, KeyError) as e:
        return None, "profiles/ncu_traffic.json unreadable: %s" % e


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
def cpu_reference_rate(cur, ref, every, threads, steps=1, warmup=0, want_results=False, pred_q=None):
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
    jobs, cands, cu_index = [], 0, []
    off = level_slices()
    if kind == "port":
        every = max(every, 20)   # the scalar oracle port: a bounded sample (about 2,150 CUs)
    for level in range(5):
        s = 8 << level
        nx = WIDTH // s
        l, r, t, b = windows(level, None if pred_q is None else pred_q[off[level]:off[level + 1]])
        for i in range(0, len(l), every):
            cu_index.append(off[level] + i)
            x, y = (i % nx) * s, (i // nx) * s
            pq = (0, 0) if pred_q is None else (int(pred_q[off[level] + i][0]), int(pred_q[off[level] + i][1]))
            jobs.append(B.make_job(cur, refp, stride, (margin + y) * stride + margin + x, s, s,
                                   (int(l[i]), int(r[i]), int(t[i]), int(b[i])), pq, 0, 0, 10, 1, 0, 1, LAMBDA,
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
    if want_results:
        # (CU index in the frame API's level-major order, (mvQx, mvQy, intX, intY, intSad, fracCost))
        out = [(cu_index[k], (4 * r.mvX + 2 * r.halfX + r.qterX, 4 * r.mvY + 2 * r.halfY + r.qterY, r.mvX, r.mvY,
                              int(r.intSad), int(r.fracCost))) for k, r in enumerate(res)]
        return cands / sec, sec, desc, kind, threads, out
    return cands / sec, sec, desc, kind, threads


def host_pair0():
    from vtm_b200.synth import make_pair
    cur, ref, _ = make_pair(0, WIDTH, HEIGHT)
    return cur, ref


WORKLOAD = ("config4: %d synthetic 1080p 10-bit frame pairs (pair p on rank p %% world), every grid-aligned square CU "
            "8..128 (43,020 per pair), full search SR=64 + half/quarter-pel SATD refinement, zero predictors (run A)" % TOTAL_PAIRS)


def shared_config():
    """The `config` object both arms print: names the workload, nothing run-specific."""
    return {"workload": WORKLOAD, "search_range": SR, "pairs": TOTAL_PAIRS, "cus_per_pair": 43020,
            "l2": "a step's inputs (32 pairs: 411 MB of planes, 4.2 GB of SAD surfaces) exceed the 126 MB L2 and steps cycle "
                  "through distinct pairs"}


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cur, ref = host_pair0()
    rate, sec, desc, kind, threads = cpu_reference_rate(cur, ref, args.cpu_every, threads, args.steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16 samples, int32 SAD/SATD arithmetic", "data": "synthetic",
            "config": shared_config(),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---- this repo's arm ------------------------------------------------------------------------------------------------
def config5_sample(ms, torch, stream, dev, hbm_peak):
    """BASELINE config 5 inside the bench run (rank 0, N = 1, after the timed legs): the table-level batch kernels on
    HBM-resident batches of 64x64 and 128x128 blocks (256 MB per operand, beyond L2), CUDA-event timed, a sample of the
    outputs compared with the oracle port.  GB/s = algorithmic bytes (operand blocks in, with the tap halo for filters, results
    out) / kernel time.  The full 36-shape sweep against the reference's AVX2 entries is microbench.py."""
    from oracle import bindings as B
    O = B.oracle()
    out = []
    gen = torch.Generator(device=dev)
    gen.manual_seed(5005)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    for (w, h) in ((64, 64), (128, 128)):
        n = (1 << 27) // (w * h)
        org = torch.randint(0, 1024, (n, h, w), device=dev, dtype=torch.int16, generator=gen)
        cur = torch.randint(0, 1024, (n, h, w), device=dev, dtype=torch.int16, generator=gen)
        res = torch.zeros(n, dtype=torch.int64, device=dev)
        sample = list(range(0, n, max(1, n // 16)))[:16]
        h_org, h_cur = org[sample].cpu().numpy(), cur[sample].cpu().numpy()
        for kind, name in ((0, "SAD"), (1, "SATD")):
            t = timed(lambda: ms.dist_batch(kind, org.data_ptr(), w, w * h, cur.data_ptr(), w, w * h, w, h, 0, n, res.data_ptr()))
            got = res[sample].cpu().numpy()
            f = O.vo_sad if kind == 0 else O.vo_satd
            want = [f(B.ptr(h_org[i]), w, B.ptr(h_cur[i]), w, w, h, 0) if kind == 0 else f(B.ptr(h_org[i]), w, B.ptr(h_cur[i]), w, w, h)
                    for i in range(len(sample))]
            nbytes = n * (2 * w * h * 2 + 8)
            out.append({"op": name, "w": w, "h": h, "blocks": n, "gbs": nbytes / t / 1e9, "hbm_frac": nbytes / t / 1e9 / hbm_peak,
                        "equal": bool([int(x) for x in got] == [int(x) for x in want])})
        del cur, res
        src = torch.randint(0, 1024, (n, h + 8, w + 8), device=dev, dtype=torch.int16, generator=gen)
        dst = torch.zeros((n, h, w), dtype=torch.int16, device=dev)
        off, ss = 4 * (w + 8) + 4, w + 8
        h_src = src[sample].cpu().numpy()
        for vert, last, name in ((0, 0, "luma 8-tap horizontal (first stage)"), (1, 1, "luma 8-tap vertical (single stage)")):
            frac = 5
            t = timed(lambda: ms.interp_batch(0, vert, src.data_ptr() + 2 * off, ss, ss * (h + 8), dst.data_ptr(), w, w * h, w, h,
                                              frac, 1, last, 10, 0, n))
            got = dst[sample].cpu().numpy()
            ok = True
            for i in range(len(sample)):
                want = np.zeros((h, w), np.int16)
                if vert:
                    O.vo_filter_ver(0, B.ptr(h_src[i], off), ss, B.ptr(want), w, w, h, frac, 1, last, 10, 0)
                else:
                    O.vo_filter_hor(0, B.ptr(h_src[i], off), ss, B.ptr(want), w, w, h, frac, last, 10, 0)
                ok = ok and bool(np.array_equal(got[i], want))
            nbytes = n * (((w + 7) * h if not vert else w * (h + 7)) * 2 + w * h * 2)
            out.append({"op": name, "w": w, "h": h, "blocks": n, "gbs": nbytes / t / 1e9, "hbm_frac": nbytes / t / 1e9 / hbm_peak,
                        "equal": ok})
        del org, src, dst
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import vtm_b200
    from vtm_b200 import FrameParams
    from vtm_b200.me import CU_RESULT_DTYPE
    from vtm_b200.peaks import int_peak
    from vtm_b200.shard import max_over_ranks, pairs_for_rank
    from vtm_b200.synth import make_pairs_torch, random_predictors

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.pairs_per_step, args.steps, args.warmup
    cands_pair, ops_pair, ncu = workload_counts()

    ms = vtm_b200.MotionSearch(local_rank)
    # a dedicated stream shared by torch (events, synthetic-data kernels) and the library (NULL would mean the
    # library's own stream, and torch's default stream handle IS NULL)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ms.set_stream(stream.cuda_stream)
    ms.set_frame_size(WIDTH, HEIGHT)

    # The rank's share of the job's pairs (pair p -> rank p % world): resident on the device as library pictures (the
    # `value` leg) and in page-locked host memory (the e2e leg).  Pair 0 is the host-generated pair the CPU leg searches,
    # so the parity check below compares the e2e leg's delivered results with the reference's on the same samples.
    ids = pairs_for_rank(args.pool, world, rank)
    pool = len(ids)
    if pool < 1:
        raise SystemExit("fewer pairs than ranks")
    host_cur, host_ref = [None] * pool, [None] * pool
    pair0 = host_pair0() if (rank == 0 and world == 1 and not args.no_cpu) else None   # CPU leg + parity: N=1 only
    for c0 in range(0, pool, 8):
        cur, ref = make_pairs_torch(ids[c0:c0 + 8], dev, WIDTH, HEIGHT)
        for i in range(cur.shape[0]):
            p = c0 + i
            if p == 0 and pair0 is not None:
                host_cur[0] = torch.from_numpy(pair0[0]).pin_memory()
                host_ref[0] = torch.from_numpy(pair0[1]).pin_memory()
                ms.upload_picture(0, pair0[0])
                ms.upload_picture(1, pair0[1])
                continue
            ms.upload_picture_device(2 * p, cur[i].data_ptr(), WIDTH, WIDTH, HEIGHT)
            ms.upload_picture_device(2 * p + 1, ref[i].data_ptr(), WIDTH, WIDTH, HEIGHT)
            host_cur[p] = cur[i].cpu().pin_memory()
            host_ref[p] = ref[i].cpu().pin_memory()
        ms.synchronize()
        del cur, ref
    prm_a = FrameParams(searchRange=SR, bitDepth=10, ctuSize=CTU, lambdaMotion=LAMBDA, predSpread=0)
    # run B of config 4: seeded random quarter-pel predictors within +-16 px, one set per pair slot of a step
    h_pred = np.stack([random_predictors(7000 + i, ncu, 16) for i in range(B)])
    d_pred = torch.from_numpy(h_pred).to(dev)
    prm_b = FrameParams(searchRange=SR, bitDepth=10, ctuSize=CTU, lambdaMotion=LAMBDA, predSpread=33)
    cands_b = int(np.mean([workload_counts(h_pred[i])[0] for i in range(min(B, 4))]))
    ops_b = int(np.mean([workload_counts(h_pred[i])[1] for i in range(min(B, 4))]))
    needed_b = int(np.mean([tree_needed_ops(h_pred[i]) for i in range(min(B, 2))]))
    d_res = torch.zeros(B * ncu * CU_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)

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
    # the kernel also runs two of its eight lanes on the FP32 pipe: the honest ceiling of that instruction mix (12 VABSDIFF : 4
    # FADD pairs per 16 pixel-candidates, variant 19 of vtmme_int_peak; profiles/r02p_mixed_pipe_peaks.md) is reported beside the
    # INT roofline SURVEY 8(d) defines
    pk_mix = int_peak(19, 1 << 16)
    mix_px = pk_mix["lane_instr_per_clk_per_sm"] / 1.25       # pixel-candidates per clock and SM
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak_ops = 2.0 * pk["lane_instr_per_clk_per_sm"] * sms * pk["sm_mhz"] * 1e6

    def timed_resident(prm, pred_ptr, steps, warm):
        """`steps` batched searches over device-resident pairs -> (ms over all steps [max over ranks], per-step kernel ms,
        launches, clocks)."""
        for s in range(warm):
            c, r = step_ids(s)
            ms.search_frames_device(c, r, prm, pred_ptr, d_res.data_ptr())
        ms.set_profiling(True)
        sampler = ClockSampler(local_rank)
        launches0 = ms.launches
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        ev0.record(stream)
        kms = []
        for s in range(warm, warm + steps):
            c, r = step_ids(s)
            ms.search_frames_device(c, r, prm, pred_ptr, d_res.data_ptr())
            kms.append(ms.frame_kernel_ms())
        ev1.record(stream)
        barrier()
        t_wall1 = time.time()
        clocks = sampler.stop(t_wall0, t_wall1)
        launches = ms.launches - launches0
        ms.set_profiling(False)
        return max_over_ranks(ev0.elapsed_time(ev1), dev), np.array(kms), launches, clocks

    run_b_headline = args.run == "B"
    prm, pred_ptr = (prm_b, d_pred.data_ptr()) if run_b_headline else (prm_a, 0)
    if run_b_headline:
        cands_pair, ops_pair = cands_b, ops_b
    elapsed_ms, kms, launches, clocks = timed_resident(prm, pred_ptr, K, W)
    value = world * B * K * cands_pair / (elapsed_ms * 1e-3)
    other = None
    if not run_b_headline:      # run B of SURVEY 8(d) beside run A: same pairs, predictor-centred windows
        b_ms, b_kms, _, _ = timed_resident(prm_b, d_pred.data_ptr(), max(2, min(K, 4)), 2)
        nb = max(2, min(K, 4))
        other = {"predictors": "seeded random quarter-pel predictors within +-16 px per CU",
                 "value": world * B * nb * cands_b / (b_ms * 1e-3), "unit": UNIT, "ms_per_step": b_ms / nb, "steps": nb,
                 "block_candidates_per_pair": cands_b,
                 "roofline_frac": B * ops_b / (float(b_kms[:, 0].mean()) * 1e-3) / peak_ops,
                 "roofline_frac_needed_ops": B * needed_b / (float(b_kms[:, 0].mean()) * 1e-3) / peak_ops,
                 "needed_ops_note": "per-CU-window ops (roofline_frac) undercount what a quad-tree search must do with "
                                    "independent predictors: an 8x8 SAD is needed on the union of the windows of the block and "
                                    "its four ancestors (tree_needed_ops: %.3g ops per pair vs %.3g)" % (needed_b, ops_b),
                 "kernel_ms": {"me_tree_sad": float(b_kms[:, 0].mean()), "me_tree_upper": float(b_kms[:, 1].mean()),
                               "me_frac_frame": float(b_kms[:, 2].mean())}}

    # ---- e2e: host buffers through the C ABI.  Every step uploads both planes of its pairs from page-locked host
    # memory (vtmme_upload_picture_async), searches them (vtmme_search_frames_device, asynchronous) and copies all
    # results back to page-locked host memory.  The three stages of consecutive steps overlap: pictures are
    # triple-buffered, results double-buffered; the timed region ends when the last result is on the host.
    e2e_steps = K
    res_bytes = B * ncu * CU_RESULT_DTYPE.itemsize
    d_res2 = [torch.zeros(res_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    h_res2 = [torch.empty(res_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    searched = [torch.cuda.Event() for _ in range(3)]
    copied = [torch.cuda.Event() for _ in range(2)]
    h_pred_pinned, d_pred_e2e = None, None
    if run_b_headline:
        h_pred_pinned = torch.from_numpy(h_pred).pin_memory()
        d_pred_e2e = torch.zeros_like(h_pred_pinned, device=dev)
    last = e2e_steps + 1          # index of the last pipelined step: it searches pool pairs 0..B-1, pair 0 in slot 0

    def e2e_pair(s, i):
        return ((s - last) * B + i) % pool

    def e2e_upload(s):
        """queue the H2D copies of step s on the library's copy stream; picture ids triple-buffered"""
        if s >= 3:
            searched[s % 3].synchronize()      # the search that last used this slot (step s-3) is done
        base = 100000 + (s % 3) * 2 * B
        for i in range(B):
            p = e2e_pair(s, i)
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
    for s in range(2, e2e_steps + 2):
        if s + 1 <= last:
            e2e_upload(s + 1)
        e2e_search(s)
    copy_stream.synchronize()                  # the last results are on the host
    ms.synchronize()
    stream.wait_stream(copy_stream)
    e1.record(stream)
    barrier()
    h_res = h_res2[last & 1].numpy().view(CU_RESULT_DTYPE).reshape(B, ncu)
    e2e_ms = max_over_ranks(e0.elapsed_time(e1), dev)
    e2e_value = world * B * e2e_steps * cands_pair / (e2e_ms * 1e-3)
    h2d = B * 2 * WIDTH * HEIGHT * 2
    d2h = B * ncu * CU_RESULT_DTYPE.itemsize

    if rank == 0:
        k1_ms = float(kms[:, 0].mean())
        k2_ms, k3_ms = float(kms[:, 1].mean()), float(kms[:, 2].mean())
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
        traffic_pair, traffic_src = ncu_traffic_per_pair()
        fops = B * frac_ops_per_pair()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16 samples, int32 SAD/SATD arithmetic", "data": "synthetic",
            "config": shared_config(),
            "run": {"predictors": "zero (run A)" if not run_b_headline else "random within +-16 px (run B)",
                    "pairs_per_step_per_gpu": B, "pool_pairs_per_gpu": pool, "block_candidates_per_pair": cands_pair,
                    "frame_pairs_per_s": world * B * K / (elapsed_ms * 1e-3)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "pinned_pool_pairs": pool},
            "gpu_launches": int(launches),
            "roofline": {"bound": "int_alu", "kernel": "me_tree_sad_kernel", "achieved": achieved / 1e12,
                         "peak": peak_ops / 1e12, "unit": "Tiop/s", "frac": achieved / peak_ops,
                         "traffic": None if traffic_pair is None else B * traffic_pair,
                         "traffic_note": "DRAM bytes per launch = pairs per launch x the per-pair figure of %s; algorithmic bytes "
                                         "per launch: %d" % (traffic_src, hbm_bytes),
                         "peak_source": "measured in this run: VABSDIFF.U32 issue rate %.1f lanes/clk/SM x %d SMs x %.0f MHz, "
                                        "fused |a-b|+c = 2 ops" % (pk["lane_instr_per_clk_per_sm"], sms, pk["sm_mhz"]),
                         "algorithmic_ops_per_launch": B * ops_pair, "kernel_ms": k1_ms,
                         "kernel_share_of_step": float(kms[:, 0].sum() / elapsed_ms),
                         "whole_step_frac": B * ops_pair / (elapsed_ms / K * 1e-3) / peak_ops,
                         "mixed_pipe_ceiling": {"what": "measured in this run: 12 VABSDIFF : 4 FADD-pair stream (the kernel's lane split), "
                                                        "ALU and FP32 pipes together", "px_cand_per_clk_per_sm": mix_px,
                                                "vs_int_pipe": mix_px / pk["lane_instr_per_clk_per_sm"],
                                                "frac": achieved / (2.0 * mix_px * sms * pk["sm_mhz"] * 1e6)},
                         "other_kernels": {
                             "me_tree_upper": {"ms": k2_ms, "bound": "hbm", "algorithmic_bytes": surf_bytes,
                                               "achieved_gbs": surf_bytes / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else None,
                                               "frac": surf_bytes / (k2_ms * 1e-3) / 1e9 / hbm_peak if k2_ms > 0 else None},
                             "me_frac_frame": {"ms": k3_ms, "bound": "int_alu", "algorithmic_ops": fops,
                                               "achieved_tiops": fops / (k3_ms * 1e-3) / 1e12,
                                               "frac": fops / (k3_ms * 1e-3) / peak_ops}},
                         "hbm": {"achieved_gbs": hbm_bytes / (k1_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "frac": hbm_bytes / (k1_ms * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s"}},
        }
        if other is not None:
            line["run_b"] = other
        if pair0 is not None:
            every = args.cpu_every
            rate, sec, desc, kind, threads, want = cpu_reference_rate(pair0[0], pair0[1], every, os.cpu_count() or 1,
                                                                      want_results=True)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc,
                                    "seconds": sec}
            # parity of what the timed e2e leg delivered for pair 0 (slot 0 of its last step) with the CPU leg's results
            # for the same host samples: every CU the CPU leg searched, all six result fields
            got0 = h_res[0]
            bad = [i for i, t in want if (int(got0["mvQx"][i]), int(got0["mvQy"][i]), int(got0["intX"][i]),
                                          int(got0["intY"][i]), int(got0["intSad"][i]), int(got0["fracCost"][i])) != t]
            line["parity"] = {"cus": len(want), "equal": not bad, "mismatches": len(bad), "against": kind,
                              "what": "e2e leg's results for host pair 0 vs the CPU leg's (MV quarter-pel, integer MV, SAD, "
                                      "refined cost)"}
            if run_b_headline:
                line["parity"] = {"cus": 0, "equal": None, "what": "run B headline: the CPU leg searches zero predictors; see "
                                                                   "tests/test_gpu_fullsize.py"}
            if kind == "reference":   # the single-thread number next to the all-cores one (sparser sample)
                r1, s1, d1, _, _ = cpu_reference_rate(pair0[0], pair0[1], 8 * every, 1)
                line["cpu_baseline"]["single_core"] = {"value": r1, "unit": UNIT, "cores": 1, "sample": d1, "seconds": s1}
        if pair0 is not None:
            line["config5"] = {"what": "table-level batch kernels (SURVEY config 5), HBM-resident batches, sample of outputs vs the "
                                       "oracle; full sweep vs the reference's AVX2 entries: profiles/*_microbench.md",
                               "hbm_peak_gbs": hbm_peak, "rows": config5_sample(ms, torch, stream, dev, hbm_peak)}
        emit(line)
        if line.get("parity", {}).get("equal") is False:
            sys.stderr.write("PARITY FAILURE: %d of %d CUs differ from the %s\n" % (len(bad), len(want), kind))
            ms.close()
            sys.exit(3)
        if any(not r["equal"] for r in line.get("config5", {}).get("rows", [])):
            sys.stderr.write("PARITY FAILURE: table-level kernels differ from the oracle (config5 rows)\n")
            ms.close()
            sys.exit(3)
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
    ap.add_argument("--pool", type=int, default=TOTAL_PAIRS, help="distinct frame pairs of the whole job (sharded p %% world)")
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
