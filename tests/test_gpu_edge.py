"""GPU: edge cases of the C ABI (degenerate and maximum windows, picture borders, tiny pictures, batches, 8-bit,
error codes) and size-independent properties at the BASELINE size (1080p, SR=64)."""
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


def _oracle_job(L, cur, refp, W, x, y, w, h, win, pq, imv=0, ssm=0, bd=10, had=1, alt=0, frac=1, lam=31.33):
    stride = refp.shape[1]
    j = B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, win, pq, imv, ssm, bd, had, alt, frac, lam,
                   org_off=y * W + x, org_stride=W)
    r = B.Result()
    L.vo_search(C.byref(j), C.byref(r), 0)
    return r.tuple()


def test_job_degenerate_border_and_max_windows(ms, oracle_lib):
    from vtm_b200 import Job
    rng = np.random.default_rng(90)
    W, H = 448, 320
    ref = rng.integers(0, 1024, (H + 2, W + 2)).astype(np.int32)
    ref = ((ref[:-2, :-2] + ref[1:-1, 1:-1] + ref[2:, 2:] + ref[:-2, 2:]) // 4).astype(np.int16)
    cur = np.ascontiguousarray(np.clip(np.roll(ref, (-4, 6), (0, 1)).astype(np.int32) + rng.integers(-6, 7, ref.shape), 0, 1023).astype(np.int16))
    refp = pad_plane(ref)
    ms.upload_picture(20, cur)
    ms.upload_picture(21, refp, MARGIN)
    cases = []
    # 1x1, 1xN and Nx1 windows
    for win in [(3, 3, -2, -2), (-5, 7, 0, 0), (1, 1, -9, 4)]:
        cases.append((64, 48, 16, 16, win, (10, -7), 0, 0))
    # windows clipped at every picture corner (xSetSearchRange with predictors pointing far outside)
    for (x, y, pq) in [(0, 0, (-4000, -4000)), (W - 32, H - 32, (4000, 4000)), (0, H - 16, (-600, 900)), (W - 64, 0, (3000, -3000))]:
        w = h = 32 if x != 0 or y != 0 else 16
        w, h = min(w, W - x), min(h, H - y)
        win = oracle_window(oracle_lib, pq, x, y, W, H, 24)
        cases.append((x, y, w, h, win, pq, 0, 0))
    # the largest block at the BASELINE search range, and a 257x257 window (SearchRange 128)
    cases.append((128, 96, 128, 128, oracle_window(oracle_lib, (0, 0), 128, 96, W, H, 64), (0, 0), 0, 2))
    cases.append((192, 128, 16, 16, oracle_window(oracle_lib, (8, -8), 192, 128, W, H, 128), (8, -8), 0, 0))
    cases.append((64, 64, 128, 64, oracle_window(oracle_lib, (-20, 12), 64, 64, W, H, 32), (-20, 12), 1, 2))
    jobs, want = [], []
    for (x, y, w, h, win, pq, imv, ssm) in cases:
        ss = oracle_lib.vo_subshift(ssm, w, h)
        jobs.append(Job(20, 21, x, y, w, h, win, pq, imv, ss, 10, 1, 1 if imv == 1 else 0, 1, 23.5))
        want.append(_oracle_job(oracle_lib, cur, refp, W, x, y, w, h, win, pq, imv, ssm, 10, 1, 1 if imv == 1 else 0, 1, 23.5))
    got = ms.search(jobs)
    bad = [(cases[i], got[i], want[i]) for i in range(len(jobs)) if got[i] != want[i]]
    assert not bad, bad[:2]


def test_job_8bit_and_sad_refinement(ms, oracle_lib):
    from vtm_b200 import Job
    rng = np.random.default_rng(91)
    W, H = 128, 96
    ref = rng.integers(0, 256, (H, W), dtype=np.int16)
    cur = np.ascontiguousarray(np.clip(np.roll(ref, (2, -3), (0, 1)) + rng.integers(-2, 3, ref.shape), 0, 255).astype(np.int16))
    refp = pad_plane(ref)
    ms.upload_picture(22, cur)
    ms.upload_picture(23, refp, MARGIN)
    jobs, want = [], []
    for (w, h, had) in [(8, 8, 1), (16, 16, 0), (32, 16, 1), (4, 8, 0), (8, 4, 1)]:
        win = oracle_window(oracle_lib, (4, 4), 32, 32, W, H, 12)
        jobs.append(Job(22, 23, 32, 32, w, h, win, (4, 4), 0, 0, 8, had, 0, 1, 9.75))
        want.append(_oracle_job(oracle_lib, cur, refp, W, 32, 32, w, h, win, (4, 4), 0, 0, 8, had, 0, 1, 9.75))
    assert ms.search(jobs) == want


@pytest.mark.parametrize("w,h", [(8, 8), (24, 16), (40, 56), (136, 72)])
def test_frame_tiny_and_ragged_pictures(ms, oracle_lib, w, h):
    """Pictures smaller than a region / CTU and not multiples of 16/32/128: the CU set shrinks level by level."""
    from vtm_b200 import FrameParams
    from vtm_b200.synth import make_pair
    cur, ref, _ = make_pair(40 + w, w, h, max_global=5, max_local=5, n_rects=0, sigma=3.0)
    refp = pad_plane(ref)
    ms.upload_picture(30, cur)
    ms.upload_picture(31, refp, MARGIN)
    ncu = ms.set_frame_size(w, h)
    got = ms.search_frames([30], [31], FrameParams(searchRange=8, lambdaMotion=17.0))
    want = oracle_frame_search(oracle_lib, cur, refp, MARGIN, 8, 17.0)
    assert ncu == len(want) and ncu >= 1
    assert [gpu_tuple(got[0][i]) for i in range(ncu)] == [want[i] for i in range(ncu)]


@pytest.mark.parametrize("w,h,sr,fast_search", [(8, 8, 8, 1), (40, 56, 16, 3), (136, 72, 128, 1), (264, 136, 200, 3)])
def test_frame_tz_tiny_ragged_and_wide_ranges(ms, oracle_lib, w, h, sr, fast_search):
    """TZ frame search on pictures smaller than a CTU / not multiples of the CU sizes (levels without any CU are
    skipped), with search ranges up to 200 (diamond distances beyond 128, windows clipped on every side), random
    predictors far outside the picture, two pairs per call."""
    from tests.helpers import oracle_frame_tz
    from vtm_b200 import FrameParams
    from vtm_b200.synth import make_pair, random_predictors
    pairs = [make_pair(90 + w + k, w, h, max_global=6, max_local=9, n_rects=1, sigma=3.0) for k in range(2)]
    refps = [pad_plane(p[1]) for p in pairs]
    for k in range(2):
        ms.upload_picture(40 + 2 * k, pairs[k][0])
        ms.upload_picture(41 + 2 * k, refps[k], MARGIN)
    ncu = ms.set_frame_size(w, h)
    pred = np.stack([random_predictors(11 + k, ncu, 300) for k in range(2)])   # up to +-300 px: clipped by clipMv
    got = ms.search_frames([40, 42], [41, 43], FrameParams(searchRange=sr, lambdaMotion=23.5, fastSearch=fast_search), pred)
    for k in range(2):
        want = oracle_frame_tz(oracle_lib, pairs[k][0], refps[k], MARGIN, sr, 23.5, pred[k], fast_search, 1, 0)
        assert [gpu_tuple(got[k][i]) for i in range(ncu)] == [want[i] for i in range(ncu)]


def test_job_tz_signed_pattern_and_amvr(ms, oracle_lib):
    """Per-call TZ jobs with a bi-pred style pattern (2*org - otherPred, outside the sample range) followed by the
    integer AMVR refinement: xTZSearch + xPatternSearchIntRefine in one call."""
    from tests.helpers import int_refine_case, tz_case
    from vtm_b200 import Amvr, Job, TzSearch
    rng = np.random.default_rng(91)
    W, H = 160, 128
    ref = np.clip(np.rint(rng.normal(512, 200, (H, W))), 0, 1023).astype(np.int16)
    cur = np.ascontiguousarray(np.roll(ref, (1, 2), (0, 1)))
    refp = pad_plane(ref)
    ms.upload_picture(44, cur)
    ms.upload_picture(45, refp, MARGIN)
    stride = refp.shape[1]
    jobs, want, keep = [], [], []
    for (w, h) in [(8, 8), (16, 32), (32, 32), (64, 16), (128, 64), (4, 16)]:
        for imv in (1, 2):
            x, y = int(rng.integers(0, (W - w) // 4 + 1)) * 4, int(rng.integers(0, (H - h) // 4 + 1)) * 4
            org = (2 * cur[y:y + h, x:x + w].astype(np.int32) - rng.integers(0, 1024, (h, w))).astype(np.int16)
            org = np.ascontiguousarray(org)
            keep.append(org)
            io = int_refine_case(rng, imv, x, y, w, h, W, H, max_pel=6)
            pred16 = (io.candX[io.mvpIdx], io.candY[io.mvpIdx])
            pq = tuple((v + 1) >> 2 if v >= 0 else (v + 2) >> 2 for v in pred16)
            t = tz_case(rng, x, y, W, H, 32, 0, 0, max_pel=10)
            t.startX, t.startY = pred16
            tz = TzSearch(pred16, 32, W, H, tuple((t.seedX[i], t.seedY[i]) for i in range(t.nSeeds)), None, 0, 0, 1)
            t.hasInt2Nx2N = 0
            am = Amvr(imv, ((io.candX[0], io.candY[0]), (io.candX[1], io.candY[1])), io.numCand, io.mvpIdx,
                      (io.mvpIdxBits[0], io.mvpIdxBits[1]), io.bits, W, H, io.fWeight)
            jobs.append(Job(44, 45, x, y, w, h, (0, 0, 0, 0), pq, imv << 1, 0, 10, 1, 0, 2, 19.0, org, am, tz))
            oj = B.make_job(org, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, (0, 0, 0, 0), pq, imv << 1, 0, 10, 1,
                            0, 0, 19.0)
            mx, my, sad = C.c_int(), C.c_int(), C.c_uint64()
            oracle_lib.vo_tz_search(C.byref(oj), C.byref(t), C.byref(mx), C.byref(my), C.byref(sad), None)
            io.mvX, io.mvY = mx.value * 16, my.value * 16
            oracle_lib.vo_int_refine(C.byref(oj), C.byref(io))
            want.append((mx.value, my.value, sad.value) + io.tuple())
    pick = lambda t: t[:3] + t[8:]
    assert [pick(t) for t in ms.search(jobs)] == want
    assert [pick(ms.search([j])[0]) for j in jobs] == want


def test_frame_batch_is_independent_per_pair(ms, oracle_lib):
    """A batch of three different pairs gives, pair by pair, what each pair gives alone (and what the oracle gives);
    running it twice gives identical results."""
    from vtm_b200 import FrameParams
    from vtm_b200.synth import make_pair
    w, h = 160, 96
    pairs = [make_pair(60 + i, w, h, max_global=6, max_local=8, n_rects=2, sigma=4.0)[:2] for i in range(3)]
    for i, (cur, ref) in enumerate(pairs):
        ms.upload_picture(40 + 2 * i, cur)
        ms.upload_picture(41 + 2 * i, pad_plane(ref), MARGIN)
    ncu = ms.set_frame_size(w, h)
    prm = FrameParams(searchRange=10, lambdaMotion=31.33)
    batch = ms.search_frames([40, 42, 44], [41, 43, 45], prm)
    again = ms.search_frames([40, 42, 44], [41, 43, 45], prm)
    assert np.array_equal(batch, again)
    for i, (cur, ref) in enumerate(pairs):
        alone = ms.search_frames([40 + 2 * i], [41 + 2 * i], prm)
        assert np.array_equal(alone[0], batch[i])
    want = oracle_frame_search(oracle_lib, pairs[1][0], pad_plane(pairs[1][1]), MARGIN, 10, 31.33)
    assert [gpu_tuple(batch[1][i]) for i in range(ncu)] == [want[i] for i in range(ncu)]


def test_error_codes(ms):
    import vtm_b200
    from vtm_b200 import FrameParams, Job
    cur = np.zeros((32, 32), np.int16)
    ms.upload_picture(50, cur)
    with pytest.raises(vtm_b200.VtmmeError, match="NOPIC"):
        ms.search([Job(50, 999, 0, 0, 8, 8, (-2, 2, -2, 2), (0, 0))])
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):
        ms.search([Job(50, 50, 0, 0, 12, 8, (-2, 2, -2, 2), (0, 0))])           # width not a power of two
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):
        ms.search([Job(50, 50, 0, 0, 8, 8, (3, 2, -2, 2), (0, 0))])             # empty window
    with pytest.raises(vtm_b200.VtmmeError, match="RANGE"):
        ms.search([Job(50, 50, 0, 0, 8, 8, (-900, 2, -2, 2), (0, 0))])          # leaves the padded picture
    ms.set_frame_size(32, 32)
    pred = np.zeros((1, ms._ncu, 2), np.int16)
    pred[0, 0] = (120, 0)                                                       # 30 px spread, declared 0
    with pytest.raises(vtm_b200.VtmmeError, match="RANGE"):
        ms.search_frames([50], [50], FrameParams(searchRange=8, predSpread=0), pred)
    with pytest.raises(vtm_b200.VtmmeError, match="NOPIC"):
        ms.release_picture(12345)
    # TZ search / AMVR refinement descriptors
    from vtm_b200 import Amvr, TzSearch
    tz = TzSearch((0, 0), 16, 32, 32)
    good = Job(50, 50, 8, 8, 8, 8, (-2, 2, -2, 2), (0, 0), tz=tz)
    assert len(ms.search([good])) == 1
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):                        # TZ and full-search jobs in one call
        ms.search([good, Job(50, 50, 8, 8, 8, 8, (-2, 2, -2, 2), (0, 0))])
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):                        # clip rectangle is not the picture's
        ms.search([Job(50, 50, 8, 8, 8, 8, (-2, 2, -2, 2), (0, 0), tz=TzSearch((0, 0), 16, 64, 32))])
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):                        # fracMode 2 without its state
        ms.search([Job(50, 50, 8, 8, 8, 8, (-2, 2, -2, 2), (0, 0), imvShift=2, fracMode=2)])
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):                        # quarter-pel CUs have no integer refinement
        ms.search([Job(50, 50, 8, 8, 8, 8, (-2, 2, -2, 2), (0, 0), imvShift=2, fracMode=2,
                       amvr=Amvr(0, ((0, 0), (0, 0)), 1, 0, (0, 0), 3, 32, 32))])
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):
        ms.search_frames([50], [50], FrameParams(searchRange=8, subShiftMode=1))   # the staged SAD belongs to fastSearch 2
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):
        ms.search_frames([50], [50], FrameParams(searchRange=200, fastSearch=2))   # selective: range <= 128
    sel = Job(50, 50, 8, 8, 8, 8, (0, 0, 0, 0), (0, 0), subShift=1, tz=TzSearch((0, 0), 16, 32, 32, selective=1, stagedSad=1))
    assert len(ms.search([sel])) == 1
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):                        # window index of the exhaustive branch
        ms.search([Job(50, 50, 8, 8, 8, 8, (0, 0, 0, 0), (0, 0), tz=TzSearch((0, 0), 200, 32, 32, selective=1))])
    with pytest.raises(vtm_b200.VtmmeError, match="ARG"):                        # staged SAD needs the mode-1 sub-shift
        ms.search([Job(50, 50, 8, 8, 8, 8, (0, 0, 0, 0), (0, 0), subShift=0, tz=TzSearch((0, 0), 16, 32, 32, stagedSad=1))])


def test_full_size_properties(ms, oracle_lib):
    """1080p, SR=64 (the BASELINE size): (a) a noise-free global integer pan is found exactly by every CU whose
    window contains it, with SAD 0; (b) a seeded sample of CUs of a noisy pair equals the oracle; (c) the search is
    idempotent."""
    from vtm_b200 import FrameParams
    from vtm_b200.synth import make_pair
    W, H = 1920, 1080
    rng = np.random.default_rng(95)
    canvas = rng.integers(0, 1024, (H + 256, W + 256)).astype(np.int16)
    gx, gy = 37, -22
    ref = np.ascontiguousarray(canvas[128:128 + H, 128:128 + W])
    cur = np.ascontiguousarray(canvas[128 + gy:128 + gy + H, 128 + gx:128 + gx + W])
    ms.upload_picture(60, cur)
    ms.upload_picture(61, ref)                      # border replicated on the device
    ncu = ms.set_frame_size(W, H)
    prm = FrameParams(searchRange=64, lambdaMotion=31.33, fracMode=0)
    res = ms.search_frames([60], [61], prm)[0]
    # CUs far enough from the picture border see true (non-replicated) reference samples at (gx, gy)
    off, checked = ms._off, 0
    for l in range(5):
        s = 8 << l
        nx = W // s
        for i in range(off[l], off[l + 1], 7):
            x, y = ((i - off[l]) % nx) * s, ((i - off[l]) // nx) * s
            if x + gx >= 0 and y + gy >= 0 and x + gx + s <= W and y + gy + s <= H:
                assert (int(res["intX"][i]), int(res["intY"][i]), int(res["intSad"][i])) == (gx, gy, 0), (l, x, y)
                checked += 1
    assert checked > 4000
    # (b) noisy synthetic pair, sample of CUs incl. the largest ones, vs oracle
    cur2, ref2, _ = make_pair(3, W, H)
    refp2 = pad_plane(ref2)
    ms.upload_picture(62, cur2)
    ms.upload_picture(63, refp2, MARGIN)
    prm2 = FrameParams(searchRange=64, lambdaMotion=31.33)
    r1 = ms.search_frames([62], [63], prm2)[0]
    sample = set(int(v) for v in rng.integers(0, off[1], 40)) | set(int(v) for v in rng.integers(off[1], off[2], 16))
    sample |= set(int(v) for v in rng.integers(off[2], off[3], 8)) | {off[3], off[3] + 211, off[4] - 1, off[4], off[4] + 59, off[5] - 1}
    want = oracle_frame_search(oracle_lib, cur2, refp2, MARGIN, 64, 31.33, only=sample)
    bad = [(i, gpu_tuple(r1[i]), want[i]) for i in sorted(sample) if gpu_tuple(r1[i]) != want[i]]
    assert not bad, bad[:3]
    # (c) idempotent
    r2 = ms.search_frames([62], [63], prm2)[0]
    assert np.array_equal(r1, r2)


def test_async_upload_pipeline(ms, oracle_lib):
    """vtmme_upload_picture_async: searches wait on the device for the pictures they name; results equal the
    synchronous path, also when an upload is queued while a search on other pictures is in flight."""
    import torch
    from vtm_b200 import FrameParams
    from vtm_b200.synth import make_pair
    w, h = 192, 128
    prm = FrameParams(searchRange=12, lambdaMotion=31.33)
    ms.set_frame_size(w, h)
    pairs = [make_pair(70 + i, w, h, max_global=8, max_local=8, n_rects=2, sigma=4.0)[:2] for i in range(2)]
    want = []
    for i, (cur, ref) in enumerate(pairs):
        ms.upload_picture(80, cur)
        ms.upload_picture(81, ref)
        want.append(ms.search_frames([80], [81], prm).copy())
    pinned = [[torch.from_numpy(a).pin_memory() for a in pr] for pr in pairs]
    # pipeline: upload pair 0, then queue pair 1's upload before searching pair 0
    ms.upload_picture_async(90, pinned[0][0].data_ptr(), w, w, h)
    ms.upload_picture_async(91, pinned[0][1].data_ptr(), w, w, h)
    ms.upload_picture_async(92, pinned[1][0].data_ptr(), w, w, h)
    ms.upload_picture_async(93, pinned[1][1].data_ptr(), w, w, h)
    got0 = ms.search_frames([90], [91], prm)
    # re-use the ids of pair 0 for pair 1 while nothing is pending on them
    ms.upload_picture_async(90, pinned[1][0].data_ptr(), w, w, h)
    ms.upload_picture_async(91, pinned[1][1].data_ptr(), w, w, h)
    got1 = ms.search_frames([92], [93], prm)
    got1b = ms.search_frames([90], [91], prm)
    ms.synchronize()
    assert np.array_equal(got0, want[0])
    assert np.array_equal(got1, want[1])
    assert np.array_equal(got1b, want[1])
