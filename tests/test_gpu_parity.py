"""GPU parity tests: the CUDA path, called through the C ABI (vtm_b200 -> libvtmme.so), against the CPU oracle
on the same seeded inputs.  Bit-exact: integer MVs, SADs, fractional MVs and costs must all be identical."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bindings as B  # noqa: E402
from tests.helpers import MARGIN, gpu_tuple, oracle_frame_search, oracle_window, pad_plane  # noqa: E402


@pytest.fixture(scope="module")
def ms():
    import vtm_b200
    m = vtm_b200.MotionSearch(0)
    yield m
    m.close()


def _frame_case(ms, oracle_lib, w, h, sr, seed, spread, lam=31.33, frac=1, use_had=1):
    from vtm_b200 import FrameParams
    from vtm_b200.synth import make_pair, random_predictors
    cur, ref, _ = make_pair(seed, w, h, max_global=min(sr - 2, 12), max_local=min(sr - 2, 14), n_rects=3, sigma=4.0)
    refp = pad_plane(ref)
    ms.upload_picture(1, cur)
    ms.upload_picture(2, refp, MARGIN)
    ncu = ms.set_frame_size(w, h)
    pred = random_predictors(seed, ncu, spread) if spread else None
    prm = FrameParams(searchRange=sr, predSpread=2 * spread + 1 if spread else 0, lambdaMotion=lam, fracMode=frac,
                      useHad=use_had)
    got = ms.search_frames([1], [2], prm, None if pred is None else pred[None])
    want = oracle_frame_search(oracle_lib, cur, refp, MARGIN, sr, lam, pred, frac, use_had)
    bad = [(i, gpu_tuple(got[0][i]), want[i]) for i in range(ncu) if gpu_tuple(got[0][i]) != want[i]]
    assert not bad, "%d of %d CUs differ, first: %s" % (len(bad), ncu, bad[:3])


def test_frame_search_zero_pred_small(ms, oracle_lib):
    _frame_case(ms, oracle_lib, 256, 128, 16, 1, 0)


def test_frame_search_partial_regions(ms, oracle_lib):
    # 200x152: neither a multiple of 32 nor of 128 -> partial regions / CTUs, border-clipped windows
    _frame_case(ms, oracle_lib, 200, 152, 24, 2, 0, lam=12.0)


def test_frame_search_random_predictors(ms, oracle_lib):
    _frame_case(ms, oracle_lib, 256, 192, 16, 3, 8)


def test_frame_search_integer_only_and_sad_frac(ms, oracle_lib):
    _frame_case(ms, oracle_lib, 128, 128, 12, 4, 0, frac=0)
    _frame_case(ms, oracle_lib, 128, 128, 12, 5, 4, use_had=0)


def test_frame_search_row_subsampling(ms, oracle_lib):
    """subShiftMode 2 (FEN=1, what the CTC encoder configs use): 16x16..64x64 CUs use 2 * SAD(even rows), 8x8 and
    128x128 every row (RdCost.cpp:310-316) — zero and random predictors, partial regions."""
    from vtm_b200 import FrameParams
    from vtm_b200.synth import make_pair, random_predictors
    for (w, h, sr, seed, spread) in [(256, 128, 16, 11, 0), (200, 152, 20, 12, 0), (256, 256, 12, 13, 6)]:
        cur, ref, _ = make_pair(seed, w, h, max_global=10, max_local=12, n_rects=3, sigma=4.0)
        refp = pad_plane(ref)
        ms.upload_picture(1, cur)
        ms.upload_picture(2, refp, MARGIN)
        ncu = ms.set_frame_size(w, h)
        pred = random_predictors(seed, ncu, spread) if spread else None
        prm = FrameParams(searchRange=sr, predSpread=2 * spread + 1 if spread else 0, lambdaMotion=27.0, subShiftMode=2)
        got = ms.search_frames([1], [2], prm, None if pred is None else pred[None])
        want = oracle_frame_search(oracle_lib, cur, refp, MARGIN, sr, 27.0, pred, sub_shift_mode=2)
        bad = [(i, gpu_tuple(got[0][i]), want[i]) for i in range(ncu) if gpu_tuple(got[0][i]) != want[i]]
        assert not bad, "%dx%d: %d of %d CUs differ, first: %s" % (w, h, len(bad), ncu, bad[:3])


def test_frame_search_8bit_and_every_level(ms, oracle_lib):
    """8-bit samples (headroom 6: the first filter stage does not shift) on a picture with 128x128 CUs, and a second
    pair in the same call (the refinement kernels index pictures, keys and results per pair)."""
    from vtm_b200 import FrameParams
    from vtm_b200.synth import make_pair, random_predictors
    w, h, sr = 256, 136, 10
    pairs = [make_pair(21 + i, w, h, max_global=7, max_local=8, n_rects=3, sigma=3.0, bit_depth=8) for i in range(2)]
    for i, (cur, ref, _) in enumerate(pairs):
        ms.upload_picture(10 + 2 * i, cur)
        ms.upload_picture(11 + 2 * i, pad_plane(ref), MARGIN)
    ncu = ms.set_frame_size(w, h)
    pred = np.stack([random_predictors(30 + i, ncu, 5) for i in range(2)])
    prm = FrameParams(searchRange=sr, bitDepth=8, predSpread=11, lambdaMotion=9.5)
    got = ms.search_frames([10, 12], [11, 13], prm, pred)
    for i, (cur, ref, _) in enumerate(pairs):
        want = oracle_frame_search(oracle_lib, cur, pad_plane(ref), MARGIN, sr, 9.5, pred[i], bit_depth=8)
        bad = [(k, gpu_tuple(got[i][k]), want[k]) for k in range(ncu) if gpu_tuple(got[i][k]) != want[k]]
        assert not bad, "pair %d: %d of %d CUs differ, first: %s" % (i, len(bad), ncu, bad[:3])


def test_frame_search_sr64(ms, oracle_lib):
    # the BASELINE search range on a picture the oracle finishes in seconds
    _frame_case(ms, oracle_lib, 256, 256, 64, 6, 0)


SIZES = [4, 8, 16, 32, 64, 128]


def test_job_search_all_shapes(ms, oracle_lib):
    """vtmme_search vs oracle: every (w,h) except 4x4, row sub-sampling, IMV shifts, signed (bi-pred) patterns."""
    from vtm_b200 import Job
    rng = np.random.default_rng(77)
    W, H = 320, 256
    ref = np.clip(np.rint(rng.normal(512, 180, (H, W))), 0, 1023).astype(np.int16)
    # low-pass so that sub-pel positions matter
    ref = ((ref[:-1, :-1].astype(np.int32) + ref[1:, :-1] + ref[:-1, 1:] + ref[1:, 1:]) // 4).astype(np.int16)
    H, W = ref.shape
    cur = np.clip(np.roll(ref, (3, -5), (0, 1)).astype(np.int32) + np.rint(rng.normal(0, 5, ref.shape)).astype(np.int32), 0, 1023).astype(np.int16)
    cur = np.ascontiguousarray(cur)
    refp = pad_plane(ref)
    ms.upload_picture(10, cur)
    ms.upload_picture(11, refp, MARGIN)
    stride = refp.shape[1]
    jobs, want = [], []
    keep = []
    for w in SIZES:
        for h in SIZES:
            if w == 4 and h == 4:
                continue
            for variant in range(3):
                x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
                y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
                sr = int(rng.integers(4, 20))
                pq = (int(rng.integers(-40, 41)), int(rng.integers(-40, 41)))
                win = oracle_window(oracle_lib, pq, x, y, W, H, sr)
                imv, alt = [(0, 0), (1, 1), (0, 0)][variant]
                ssm = 2 if variant == 2 else 0
                ss = oracle_lib.vo_subshift(ssm, w, h)
                lam = float(rng.uniform(4, 60))
                org = None
                if variant == 1:   # bi-pred style pattern: 2*org - otherPred, outside [0,1023]
                    org = (2 * cur[y:y + h, x:x + w].astype(np.int32) - rng.integers(0, 1024, (h, w))).astype(np.int16)
                    org = np.ascontiguousarray(org)
                    keep.append(org)
                jobs.append(Job(10, 11, x, y, w, h, win, pq, imv, ss, 10, 1, alt, 1, lam, org))
                o_arr, o_off, o_stride = (cur, y * W + x, W) if org is None else (org, 0, w)
                oj = B.make_job(o_arr, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, win, pq, imv, ssm, 10, 1,
                                alt, 1, lam, org_off=o_off, org_stride=o_stride)
                r = B.Result()
                oracle_lib.vo_search(C.byref(oj), C.byref(r), 0)
                want.append(r.tuple())
    got = ms.search(jobs)
    bad = [(i, jobs[i].w, jobs[i].h, got[i], want[i]) for i in range(len(jobs)) if got[i] != want[i]]
    assert not bad, "%d of %d jobs differ, first: %s" % (len(bad), len(jobs), bad[:3])
    # one job per call: what the in-loop encoder does (patterns up to 32x32 take the single-launch path)
    got1 = [ms.search([j])[0] for j in jobs]
    bad = [(i, jobs[i].w, jobs[i].h, got1[i], want[i]) for i in range(len(jobs)) if got1[i] != want[i]]
    assert not bad, "single calls: %d of %d jobs differ, first: %s" % (len(bad), len(jobs), bad[:3])


def test_job_search_integer_amvr(ms, oracle_lib):
    """imvShift 2/4 (FPEL / 4PEL): xPatternSearchFracDIF's integer-only branch (one SATD + rate at scale 2)."""
    from vtm_b200 import Job
    rng = np.random.default_rng(78)
    W, H = 192, 160
    ref = rng.integers(0, 1024, (H, W), dtype=np.int16)
    cur = np.ascontiguousarray(np.roll(ref, (-2, 4), (0, 1)))
    refp = pad_plane(ref)
    ms.upload_picture(12, cur)
    ms.upload_picture(13, refp, MARGIN)
    stride = refp.shape[1]
    jobs, want = [], []
    for (w, h) in [(8, 8), (16, 8), (8, 16), (32, 32), (64, 32), (4, 8)]:
        for imv in (2, 4):
            x, y = 48, 40
            pq = (int(rng.integers(-20, 21)) * 4, int(rng.integers(-20, 21)) * 4)
            win = oracle_window(oracle_lib, pq, x, y, W, H, 10)
            jobs.append(Job(12, 13, x, y, w, h, win, pq, imv, 0, 10, 1, 0, 1, 20.0))
            oj = B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, win, pq, imv, 0, 10, 1, 0, 1,
                            20.0, org_off=y * W + x, org_stride=W)
            r = B.Result()
            oracle_lib.vo_search(C.byref(oj), C.byref(r), 0)
            want.append(r.tuple())
    got = ms.search(jobs)
    assert got == want
    assert [ms.search([j])[0] for j in jobs] == want   # single-launch path of the small patterns


def test_job_search_int_refine(ms, oracle_lib):
    """fracMode 2: the integer search followed by xPatternSearchIntRefine (InterSearch.cpp:4172-4282) — FPEL and
    4PEL AMVR, SATD and SAD, one or two AMVP candidates, bi-pred weights and patterns, border clipping; batched call,
    single-launch path (patterns <= 32x32) and the multi-CTA path of larger patterns."""
    from tests.helpers import int_refine_case
    from vtm_b200 import Amvr, Job
    rng = np.random.default_rng(79)
    W, H = 256, 192
    ref = np.clip(np.rint(rng.normal(512, 200, (H, W))), 0, 1023).astype(np.int16)
    cur = np.clip(np.roll(ref, (2, -3), (0, 1)).astype(np.int32) + np.rint(rng.normal(0, 6, ref.shape)).astype(np.int32), 0, 1023)
    cur = np.ascontiguousarray(cur.astype(np.int16))
    refp = pad_plane(ref)
    ms.upload_picture(14, cur)
    ms.upload_picture(15, refp, MARGIN)
    stride = refp.shape[1]
    jobs, want, keep = [], [], []
    for w in SIZES:
        for h in SIZES:
            if w == 4 and h == 4:
                continue
            for variant in range(4):
                imv = 1 + (variant & 1)
                use_had = 0 if variant == 3 else 1
                # positions at the picture border too, so that clipMv matters for the probes
                x = [0, W - w][variant & 1] if variant >= 2 else int(rng.integers(0, (W - w) // 4 + 1)) * 4
                y = [0, H - h][(variant >> 1) & 1] if variant >= 2 else int(rng.integers(0, (H - h) // 4 + 1)) * 4
                sr = int(rng.integers(3, 12))
                lam = float(rng.uniform(4, 60))
                io = int_refine_case(rng, imv, x, y, w, h, W, H, max_pel=6)
                pred16 = (io.candX[io.mvpIdx], io.candY[io.mvpIdx])
                # xMotionEstimation: predictor INTERNAL -> QUARTER (InterSearch.cpp:3370-3372), window around it
                pq = tuple((v + 1) >> 2 if v >= 0 else (v + 2) >> 2 for v in pred16)
                win = oracle_window(oracle_lib, pq, x, y, W, H, sr)
                imv_shift = imv << 1
                org = None
                if variant == 1:
                    org = (2 * cur[y:y + h, x:x + w].astype(np.int32) - rng.integers(0, 1024, (h, w))).astype(np.int16)
                    org = np.ascontiguousarray(org)
                    keep.append(org)
                am = Amvr(imv, ((io.candX[0], io.candY[0]), (io.candX[1], io.candY[1])), io.numCand, io.mvpIdx,
                          (io.mvpIdxBits[0], io.mvpIdxBits[1]), io.bits, W, H, io.fWeight)
                jobs.append(Job(14, 15, x, y, w, h, win, pq, imv_shift, 0, 10, use_had, 0, 2, lam, org, am))
                o_arr, o_off, o_stride = (cur, y * W + x, W) if org is None else (org, 0, w)
                oj = B.make_job(o_arr, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, win, pq, imv_shift, 0, 10,
                                use_had, 0, 0, lam, org_off=o_off, org_stride=o_stride)
                r = B.Result()
                oracle_lib.vo_search(C.byref(oj), C.byref(r), 0)
                io.mvX, io.mvY = r.mvX * 16, r.mvY * 16   # rcMv.changePrecision(INT -> INTERNAL), :3488
                oracle_lib.vo_int_refine(C.byref(oj), C.byref(io))
                want.append((r.mvX, r.mvY, r.intSad) + io.tuple())
    pick = lambda t: t[:3] + t[8:]
    got = [pick(t) for t in ms.search(jobs)]
    bad = [(i, jobs[i].w, jobs[i].h, got[i], want[i]) for i in range(len(jobs)) if got[i] != want[i]]
    assert not bad, "%d of %d jobs differ, first: %s" % (len(bad), len(jobs), bad[:3])
    got1 = [pick(ms.search([j])[0]) for j in jobs]
    bad = [(i, jobs[i].w, jobs[i].h, got1[i], want[i]) for i in range(len(jobs)) if got1[i] != want[i]]
    assert not bad, "single calls: %d of %d jobs differ, first: %s" % (len(bad), len(jobs), bad[:3])


@pytest.mark.parametrize("extended,fast", [(0, 0), (1, 0), (0, 1)])
def test_job_tz_search(ms, oracle_lib, extended, fast):
    """vtmme_search with vtmme_tz: xTZSearch (InterSearch.cpp:3640-3974) as the integer search — FastSearch=1,
    FastSearch=3 (extended) and the fast re-search — followed by the fractional refinement, against the oracle
    (which is pinned on the reference's own xTZSearch): all shapes, FEN row sub-sampling, history seeds with
    duplicates, 2Nx2N integer MV, border positions, real motion so that raster scan and star refinement run."""
    from tests.helpers import tz_case
    from vtm_b200 import Job, TzSearch
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(300 + 2 * extended + fast)
    W, H = 256, 192
    cur, ref, _ = make_pair(60 + extended, W, H, max_global=20, max_local=30, n_rects=4, sigma=6.0)
    refp = pad_plane(ref)
    ms.upload_picture(16, cur)
    ms.upload_picture(17, refp, MARGIN)
    stride = refp.shape[1]
    jobs, want = [], []
    for w in SIZES:
        for h in SIZES:
            if w == 4 and h == 4:
                continue
            for rep in range(3):
                x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
                y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
                if rep == 2:
                    x, y = [0, W - w][int(rng.integers(0, 2))], [0, H - h][int(rng.integers(0, 2))]
                sr = [64, 32, 96][rep]
                t = tz_case(rng, x, y, W, H, sr, extended, fast, first_stop=int(rep != 1), max_pel=20 if rep < 2 else 160)
                pq = (int(rng.integers(-80, 81)), int(rng.integers(-80, 81)))
                ssm = [0, 2, 2][rep]
                ss = oracle_lib.vo_subshift(ssm, w, h)
                lam = 31.33 if rep else 8.5
                tz = TzSearch((t.startX, t.startY), sr, W, H, tuple((t.seedX[i], t.seedY[i]) for i in range(t.nSeeds)),
                              (t.int2Nx2NX, t.int2Nx2NY) if t.hasInt2Nx2N else None, extended, fast, t.firstSearchStop)
                jobs.append(Job(16, 17, x, y, w, h, (0, 0, 0, 0), pq, 0, ss, 10, 1, 0, 1, lam, None, None, tz))
                oj = B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, (0, 0, 0, 0), pq, 0, ssm, 10, 1,
                                0, 1, lam, org_off=y * W + x, org_stride=W)
                mx, my, sad = C.c_int(), C.c_int(), C.c_uint64()
                oracle_lib.vo_tz_search(C.byref(oj), C.byref(t), C.byref(mx), C.byref(my), C.byref(sad), None)
                hx, hy, qx, qy, cost = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
                oracle_lib.vo_frac_direct(C.byref(oj), mx.value, my.value, C.byref(hx), C.byref(hy), C.byref(qx),
                                          C.byref(qy), C.byref(cost))
                want.append((mx.value, my.value, sad.value, hx.value, hy.value, qx.value, qy.value, cost.value))
    got = ms.search(jobs)
    bad = [(i, jobs[i].w, jobs[i].h, got[i], want[i]) for i in range(len(jobs)) if got[i] != want[i]]
    assert not bad, "%d of %d jobs differ, first: %s" % (len(bad), len(jobs), bad[:3])
    got1 = [ms.search([j])[0] for j in jobs[::7]]
    assert got1 == want[::7]


@pytest.mark.parametrize("selective,fast", [(1, 0), (0, 1), (0, 0)])
def test_tz_selective_jobs(ms, oracle_lib, selective, fast):
    """vtmme_search with vtmme_tz.selective / stagedSad: xTZSearchSelective (InterSearch.cpp:3979-4170, FastSearch=2) and
    the staged SAD of subShiftMode 1 (xTZSearchHelp :340-391; also what the cached-MV re-search through xTZSearch uses
    under FastSearch=2) followed by the fractional refinement, against the oracle (pinned on the reference's own
    members): all shapes (sub-shifts 1-4), border positions, far motion (the exhaustive branch) and near motion (star
    refinement), history seeds, sub-sampling modes 1 / 2 / 0."""
    from tests.helpers import tz_case
    from vtm_b200 import Job, TzSearch
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(800 + 2 * selective + fast)
    W, H = 256, 192
    jobs, want = [], []
    for seed in range(2):
        cur, ref, _ = make_pair(90 + seed, W, H, max_global=[3, 24][seed], max_local=30, n_rects=4, sigma=6.0)
        refp = pad_plane(ref)
        ms.upload_picture(40 + 2 * seed, cur)
        ms.upload_picture(41 + 2 * seed, refp, MARGIN)
        stride = refp.shape[1]
        for w in SIZES:
            for h in SIZES:
                if w == 4 and h == 4:
                    continue
                for rep in range(3):
                    x = int(rng.integers(0, (W - w) // 4 + 1)) * 4
                    y = int(rng.integers(0, (H - h) // 4 + 1)) * 4
                    if rep == 2:
                        x, y = [0, W - w][int(rng.integers(0, 2))], [0, H - h][int(rng.integers(0, 2))]
                    sr = [64, 32, 96][rep]
                    t = tz_case(rng, x, y, W, H, sr, 0, fast, first_stop=int(rep != 1), max_pel=20 if rep < 2 else 160)
                    t.selective = selective
                    pq = (int(rng.integers(-80, 81)), int(rng.integers(-80, 81)))
                    ssm = [1, 1, 2 * selective][rep] if (selective or fast) else 1
                    ss = oracle_lib.vo_subshift(ssm, w, h)
                    lam = 31.33 if rep else 8.5
                    tz = TzSearch((t.startX, t.startY), sr, W, H, tuple((t.seedX[i], t.seedY[i]) for i in range(t.nSeeds)),
                                  (t.int2Nx2NX, t.int2Nx2NY) if t.hasInt2Nx2N else None, 0, fast, t.firstSearchStop, 128,
                                  selective, int(ssm == 1))
                    jobs.append(Job(40 + 2 * seed, 41 + 2 * seed, x, y, w, h, (0, 0, 0, 0), pq, 0, ss, 10, 1, 0, 1, lam, None,
                                    None, tz))
                    oj = B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, (0, 0, 0, 0), pq, 0, ssm, 10,
                                    1, 0, 1, lam, org_off=y * W + x, org_stride=W)
                    mx, my, sad = C.c_int(), C.c_int(), C.c_uint64()
                    oracle_lib.vo_tz_search(C.byref(oj), C.byref(t), C.byref(mx), C.byref(my), C.byref(sad), None)
                    hx, hy, qx, qy, cost = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
                    oracle_lib.vo_frac_direct(C.byref(oj), mx.value, my.value, C.byref(hx), C.byref(hy), C.byref(qx),
                                              C.byref(qy), C.byref(cost))
                    want.append((mx.value, my.value, sad.value, hx.value, hy.value, qx.value, qy.value, cost.value))
    got = ms.search(jobs)
    bad = [(i, jobs[i].w, jobs[i].h, got[i], want[i]) for i in range(len(jobs)) if got[i] != want[i]]
    assert not bad, "%d of %d jobs differ, first: %s" % (len(bad), len(jobs), bad[:3])
    got1 = [ms.search([j])[0] for j in jobs[::11]]     # the single-launch path
    assert got1 == want[::11]


@pytest.mark.parametrize("fast_search,ssm,spread", [(1, 0, 0), (1, 2, 9), (3, 0, 9), (3, 2, 0), (2, 1, 9), (2, 2, 0), (2, 0, 30)])
def test_frame_tz_search(ms, oracle_lib, fast_search, ssm, spread):
    """vtmme_search_frames with fastSearch 1 / 3: every CU's integer search is xTZSearch started at its predictor;
    fastSearch 2: xTZSearchSelective, with the staged SAD when subShiftMode is 1 (spread 30: far predictors, so that the
    exhaustive branch runs)."""
    from tests.helpers import oracle_frame_tz
    from vtm_b200 import FrameParams
    from vtm_b200.synth import make_pair, random_predictors
    w, h, sr, lam = 320, 200, 64, 31.33
    cur, ref, _ = make_pair(70 + fast_search + ssm, w, h, max_global=20, max_local=28, n_rects=4, sigma=5.0)
    refp = pad_plane(ref)
    ms.upload_picture(18, cur)
    ms.upload_picture(19, refp, MARGIN)
    ncu = ms.set_frame_size(w, h)
    pred = random_predictors(5, ncu, spread) if spread else None
    prm = FrameParams(searchRange=sr, lambdaMotion=lam, subShiftMode=ssm, fastSearch=fast_search,
                      predSpread=2 * spread + 1 if spread else 0)
    got = ms.search_frames([18], [19], prm, None if pred is None else pred[None])
    want = oracle_frame_tz(oracle_lib, cur, refp, MARGIN, sr, lam, pred, fast_search, 1, ssm)
    bad = [(i, gpu_tuple(got[0][i]), want[i]) for i in range(ncu) if gpu_tuple(got[0][i]) != want[i]]
    assert not bad, "%d of %d CUs differ, first: %s" % (len(bad), ncu, bad[:3])


def test_dist_host_all_shapes(ms, oracle_lib):
    """The DistParam-hook flavour: one block pair in host memory, SAD (subShift 0/1) and SATD, incl. 4x4."""
    rng = np.random.default_rng(79)
    for w in SIZES:
        for h in SIZES:
            org = rng.integers(0, 1024, (h, w + 3), dtype=np.int16)
            cur = np.clip(org.astype(np.int32) + np.rint(rng.normal(0, 16, org.shape)).astype(np.int32), 0, 1023).astype(np.int16)
            o, c = org[:, 1:1 + w], cur[:, 2:2 + w]
            for ss in (0, 1):
                if (h >> ss) < 1:
                    continue
                assert ms.dist_host(0, o, c, ss) == oracle_lib.vo_sad(B.ptr(org, 1), w + 3, B.ptr(cur, 2), w + 3, w, h, ss)
            assert ms.dist_host(1, o, c) == oracle_lib.vo_satd(B.ptr(org, 1), w + 3, B.ptr(cur, 2), w + 3, w, h)


def test_dist_batch_device(ms, oracle_lib):
    import torch
    rng = np.random.default_rng(80)
    n = 512
    for (w, h) in [(8, 8), (16, 16), (32, 8), (8, 32), (64, 64), (4, 16), (16, 4), (128, 128), (4, 4)]:
        org = rng.integers(0, 1024, (n, h, w), dtype=np.int16)
        cur = rng.integers(0, 1024, (n, h, w), dtype=np.int16)
        d_org, d_cur = torch.from_numpy(org).cuda(), torch.from_numpy(cur).cuda()
        out = torch.zeros(n, dtype=torch.int64, device="cuda")
        for kind, ss in ((0, 0), (0, 1), (1, 0)):
            if (h >> ss) < 1:
                continue
            ms.dist_batch(kind, d_org.data_ptr(), w, w * h, d_cur.data_ptr(), w, w * h, w, h, ss, n, out.data_ptr())
            ms.synchronize()
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            for i in range(0, n, 37):
                if kind == 0:
                    ref = oracle_lib.vo_sad(B.ptr(org[i]), w, B.ptr(cur[i]), w, w, h, ss)
                else:
                    ref = oracle_lib.vo_satd(B.ptr(org[i]), w, B.ptr(cur[i]), w, w, h)
                assert int(got[i]) == ref, (w, h, kind, ss, i)


def test_interp_host_matches_oracle(ms, oracle_lib):
    rng = np.random.default_rng(81)
    for bd in (8, 10):
        for (w, h) in [(4, 4), (4, 11), (8, 8), (9, 16), (17, 24), (64, 64), (129, 136)]:
            for frac in (0, 1, 4, 8, 12, 15):
                for alt in ((0, 1) if frac == 8 else (0,)):
                    src = rng.integers(0, 1 << bd, (h + 10, w + 10), dtype=np.int16)
                    mid = rng.integers(-8192, 8192, (h + 10, w + 10), dtype=np.int16)
                    off, ss = 4 * (w + 10) + 4, w + 10
                    for (vert, first, last, s) in [(0, 1, 0, src), (0, 1, 1, src), (1, 1, 0, src), (1, 1, 1, src),
                                                   (1, 0, 1, mid), (1, 0, 0, mid)]:
                        want = np.zeros((h, w), np.int16)
                        if vert:
                            oracle_lib.vo_filter_ver(0, B.ptr(s, off), ss, B.ptr(want), w, w, h, frac, first, last, bd, alt)
                        else:
                            oracle_lib.vo_filter_hor(0, B.ptr(s, off), ss, B.ptr(want), w, w, h, frac, last, bd, alt)
                        got = ms.interp_host(0, vert, s, off, ss, w, h, frac, first, last, bd, alt)
                        assert np.array_equal(got, want), (bd, w, h, frac, alt, vert, first, last)
    for (w, h) in [(2, 2), (4, 4), (8, 2), (16, 16), (64, 64)]:
        for frac in (0, 1, 7, 16, 31):
            src = rng.integers(0, 1024, (h + 6, w + 6), dtype=np.int16)
            off, ss = 2 * (w + 6) + 2, w + 6
            for (vert, first, last) in [(0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 1, 0)]:
                want = np.zeros((h, w), np.int16)
                if vert:
                    oracle_lib.vo_filter_ver(1, B.ptr(src, off), ss, B.ptr(want), w, w, h, frac, first, last, 10, 0)
                else:
                    oracle_lib.vo_filter_hor(1, B.ptr(src, off), ss, B.ptr(want), w, w, h, frac, last, 10, 0)
                got = ms.interp_host(1, vert, src, off, ss, w, h, frac, first, last, 10, 0)
                assert np.array_equal(got, want), ("chroma", w, h, frac, vert, first, last)


LUMA = {4: (-1, 4, -10, 58, 17, -5, 1, 0), 8: (-1, 4, -11, 40, 40, -11, 4, -1), 13: (0, 1, -4, 13, 60, -8, 3, -1)}
CHROMA = {5: (-3, 57, 12, -2), 16: (-4, 36, 36, -4)}


def test_filter_host_explicit_taps(ms, oracle_lib):
    """vtmme_filter_host (the table-entry flavour with explicit coefficients) against the oracle's dispatch."""
    rng = np.random.default_rng(82)
    w, h = 24, 16
    src = rng.integers(0, 1024, (h + 10, w + 10), dtype=np.int16)
    mid = rng.integers(-8192, 8192, (h + 10, w + 10), dtype=np.int16)
    off, ss = 4 * (w + 10) + 4, w + 10
    for comp, table, taps in ((0, LUMA, 8), (1, CHROMA, 4)):
        for frac, coeff in table.items():
            for (vert, first, last, s) in [(0, 1, 0, src), (0, 1, 1, src), (1, 0, 1, mid), (1, 1, 1, src), (1, 0, 0, mid)]:
                want = np.zeros((h, w), np.int16)
                if vert:
                    oracle_lib.vo_filter_ver(comp, B.ptr(s, off), ss, B.ptr(want), w, w, h, frac, first, last, 10, 0)
                else:
                    oracle_lib.vo_filter_hor(comp, B.ptr(s, off), ss, B.ptr(want), w, w, h, frac, last, 10, 0)
                got = ms.filter_host(taps, vert, first, last, 0, s, off, ss, w, h, coeff, 10)
                assert np.array_equal(got, want), (comp, frac, vert, first, last)
    for (first, last, s) in [(1, 0, src), (0, 1, mid)]:      # filterCopy<true,false> / <false,true>
        want = np.zeros((h, w), np.int16)
        oracle_lib.vo_filter_ver(0, B.ptr(s, off), ss, B.ptr(want), w, w, h, 0, first, last, 10, 0)
        got = ms.filter_host(8, 1, first, last, 1, s, off, ss, w, h, None, 10)
        assert np.array_equal(got, want), ("copy", first, last)
