"""Pins the C oracle (oracle/vtm_oracle.c) against the UNMODIFIED reference compiled here
(oracle/_ref/libvtmref.so = VTM 9.3 built by oracle/Makefile.ref).  CPU only."""
import ctypes as C

import numpy as np
import pytest

from oracle import bindings as B

SIZES = [4, 8, 16, 32, 64, 128]
SHAPES = [(w, h) for w in SIZES for h in SIZES]


def blocks(rng, w, h, bd=10, kind="uniform", signed_org=False):
    hi = 1 << bd
    cur = rng.integers(0, hi, (h + 3, w + 5), dtype=np.int16)
    if kind == "uniform":
        org = rng.integers(0, hi, (h, w + 2), dtype=np.int16)
    else:
        org = np.clip(cur[:h, :w + 2].astype(np.int32) + np.rint(rng.normal(0, 16, (h, w + 2))).astype(np.int32), 0, hi - 1).astype(np.int16)
    if signed_org:   # bi-pred "2*org - otherPred" range (SURVEY hard part 5)
        org = (2 * org.astype(np.int32) - rng.integers(0, hi, org.shape)).astype(np.int16)
    return np.ascontiguousarray(org), np.ascontiguousarray(cur)


@pytest.mark.ref
def test_ref_uses_simd(ref_lib):
    assert ref_lib.ref_simd_level() >= 1   # SSE4.1 or better: the SIMD tables are the ones pinned here


@pytest.mark.ref
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_subshift(oracle_lib, ref_lib, mode):
    for w, h in SHAPES:
        assert oracle_lib.vo_subshift(mode, w, h) == ref_lib.ref_subshift(mode, w, h)


@pytest.mark.ref
@pytest.mark.parametrize("kind,signed_org", [("uniform", False), ("residual", False), ("uniform", True)])
def test_sad_satd_all_shapes(oracle_lib, ref_lib, kind, signed_org):
    rng = np.random.default_rng(7)
    for w, h in SHAPES:
        for rep in range(4):
            org, cur = blocks(rng, w, h, 10, kind, signed_org)
            os_, cs = org.shape[1], cur.shape[1]
            for mode in (0, 2):
                ss = oracle_lib.vo_subshift(mode, w, h)
                a = oracle_lib.vo_sad(B.ptr(org), os_, B.ptr(cur, 1), cs, w, h, ss)
                b = ref_lib.ref_dist(B.ptr(org), os_, B.ptr(cur, 1), cs, w, h, 10, mode, 0)
                assert a == b, (w, h, mode)
            a = oracle_lib.vo_satd(B.ptr(org), os_, B.ptr(cur, 1), cs, w, h)
            b = ref_lib.ref_dist(B.ptr(org), os_, B.ptr(cur, 1), cs, w, h, 10, 0, 1)
            assert a == b, (w, h, "satd")


@pytest.mark.ref
def test_sad_satd_8bit(oracle_lib, ref_lib):
    rng = np.random.default_rng(8)
    for w, h in SHAPES:
        org, cur = blocks(rng, w, h, 8)
        os_, cs = org.shape[1], cur.shape[1]
        assert oracle_lib.vo_sad(B.ptr(org), os_, B.ptr(cur), cs, w, h, 0) == ref_lib.ref_dist(B.ptr(org), os_, B.ptr(cur), cs, w, h, 8, 0, 0)
        assert oracle_lib.vo_satd(B.ptr(org), os_, B.ptr(cur), cs, w, h) == ref_lib.ref_dist(B.ptr(org), os_, B.ptr(cur), cs, w, h, 8, 0, 1)


@pytest.mark.ref
def test_mv_bits_and_cost(oracle_lib, ref_lib):
    rng = np.random.default_rng(9)
    for _ in range(20000):
        x, y = (int(v) for v in rng.integers(-600, 600, 2))
        px, py = (int(v) for v in rng.integers(-2400, 2400, 2))
        scale = int(rng.integers(0, 3))
        imv = int(rng.choice([0, 1, 2, 4]))
        assert oracle_lib.vo_mv_bits(x, y, px, py, scale, imv) == ref_lib.ref_mv_bits(x, y, px, py, scale, imv)
    for lam in (31.33, 4.7, 57.908, 123.456789):
        for bits in range(0, 140):
            assert oracle_lib.vo_mv_cost(lam, bits) == ref_lib.ref_mv_cost(lam, bits)


def _filt_case(oracle_lib, ref_lib, rng, comp, w, h, frac, bd, first, last, alt, vertical):
    taps = 8 if comp == 0 else 4
    hi = 1 << bd
    sh, sw = h + taps + 2, w + taps + 2
    if first:
        src = rng.integers(0, hi, (sh, sw), dtype=np.int16)
    else:   # 14-bit intermediates as a first stage would produce
        src = rng.integers(-8192, 8192, (sh, sw), dtype=np.int16)
    off = (taps // 2) * sw + taps // 2
    d0 = np.zeros((h, w), np.int16)
    d1 = np.zeros((h, w), np.int16)
    if vertical:
        oracle_lib.vo_filter_ver(comp, B.ptr(src, off), sw, B.ptr(d0), w, w, h, frac, first, last, bd, alt)
        ref_lib.ref_filter_ver(comp, B.ptr(src, off), sw, B.ptr(d1), w, w, h, frac, first, last, bd, alt)
    else:
        oracle_lib.vo_filter_hor(comp, B.ptr(src, off), sw, B.ptr(d0), w, w, h, frac, last, bd, alt)
        ref_lib.ref_filter_hor(comp, B.ptr(src, off), sw, B.ptr(d1), w, w, h, frac, last, bd, alt)
    assert np.array_equal(d0, d1), (comp, w, h, frac, bd, first, last, alt, vertical)


@pytest.mark.ref
@pytest.mark.parametrize("bd", [8, 10])
def test_interpolation_luma(oracle_lib, ref_lib, bd):
    rng = np.random.default_rng(10)
    for (w, h) in [(4, 4), (4, 11), (4, 8), (8, 4), (5, 12), (9, 16), (17, 24), (16, 16), (33, 40), (64, 64), (128, 128), (129, 136)]:
        for frac in range(16):
            for alt in (0, 1):
                if alt and frac != 8:
                    continue
                _filt_case(oracle_lib, ref_lib, rng, 0, w, h, frac, bd, 1, 0, alt, False)
                _filt_case(oracle_lib, ref_lib, rng, 0, w, h, frac, bd, 1, 1, alt, False)
                _filt_case(oracle_lib, ref_lib, rng, 0, w, h, frac, bd, 1, 0, alt, True)
                _filt_case(oracle_lib, ref_lib, rng, 0, w, h, frac, bd, 1, 1, alt, True)
                _filt_case(oracle_lib, ref_lib, rng, 0, w, h, frac, bd, 0, 1, alt, True)
                _filt_case(oracle_lib, ref_lib, rng, 0, w, h, frac, bd, 0, 0, alt, True)


@pytest.mark.ref
def test_interpolation_chroma(oracle_lib, ref_lib):
    rng = np.random.default_rng(11)
    for (w, h) in [(2, 2), (4, 4), (2, 8), (8, 2), (16, 16), (64, 64), (3, 7)]:
        for frac in range(32):
            _filt_case(oracle_lib, ref_lib, rng, 1, w, h, frac, 10, 1, 0, 0, False)
            _filt_case(oracle_lib, ref_lib, rng, 1, w, h, frac, 10, 1, 1, 0, False)
            _filt_case(oracle_lib, ref_lib, rng, 2, w, h, frac, 10, 1, 1, 0, True)
            _filt_case(oracle_lib, ref_lib, rng, 2, w, h, frac, 10, 0, 1, 0, True)


def planted(rng, w, h, sr, smooth=True):
    """Reference plane with margin and an original block = displaced + noisy copy of it."""
    m = sr + 16
    H, W = h + 2 * m, w + 2 * m
    ref = rng.integers(0, 1024, (H + 2, W + 2)).astype(np.float64)
    if smooth:
        ref = (ref[:-2, :-2] + ref[:-2, 1:-1] + ref[:-2, 2:] + ref[1:-1, :-2] + ref[1:-1, 1:-1] + ref[1:-1, 2:] + ref[2:, :-2] + ref[2:, 1:-1] + ref[2:, 2:]) / 9
    else:
        ref = ref[:H, :W]
    ref = np.ascontiguousarray(np.clip(np.rint(ref), 0, 1023).astype(np.int16))
    dx, dy = (int(v) for v in rng.integers(-sr + 2, sr - 1, 2))
    org = ref[m + dy:m + dy + h, m + dx:m + dx + w].astype(np.int32)
    org = np.clip(org + np.rint(rng.normal(0, 6, org.shape)).astype(np.int32), 0, 1023).astype(np.int16)
    return ref, np.ascontiguousarray(org), m * W + m, W, (dx, dy)


@pytest.mark.ref
@pytest.mark.parametrize("w,h", [(8, 8), (16, 16), (8, 4), (4, 8), (16, 8), (8, 16), (32, 8), (8, 32), (32, 32), (64, 16), (16, 64), (64, 64), (128, 64), (128, 128)])
def test_search_matches_reference(oracle_lib, ref_lib, w, h):
    rng = np.random.default_rng(100 * w + h)
    nrep = 3 if w * h <= 1024 else 1
    for rep in range(nrep):
        sr = 24 if w * h > 4096 else 40
        ref, org, off, stride, _ = planted(rng, w, h, sr)
        for imv, alt, ssm in [(0, 0, 0), (0, 0, 2), (1, 1, 0), (2, 0, 2)]:
            pred = (int(rng.integers(-40, 40)), int(rng.integers(-40, 40)))
            win = (-sr + int(rng.integers(0, 5)), sr - int(rng.integers(0, 5)), -sr + int(rng.integers(0, 5)), sr - int(rng.integers(0, 5)))
            do_frac = 0 if imv > 1 else 1
            j = B.make_job(org, ref, stride, off, w, h, win, pred, imv, ssm, 10, 1, alt, do_frac, 31.33 + rep)
            r0, r1, r2 = B.Result(), B.Result(), B.Result()
            ref_lib.ref_search(C.byref(j), C.byref(r0))
            oracle_lib.vo_search(C.byref(j), C.byref(r1), 1)
            oracle_lib.vo_search(C.byref(j), C.byref(r2), 0)
            assert r0.tuple() == r1.tuple(), ("literal", w, h, imv, ssm)
            assert r0.tuple() == r2.tuple(), ("direct", w, h, imv, ssm)


def planted_frac(oracle_lib, rng, w, h, sr):
    """Like planted(), but the original block sits at a random quarter-pel phase of the reference."""
    ref, org, off, stride, (dx, dy) = planted(rng, w, h, sr)
    dq = (int(rng.integers(-3, 4)), int(rng.integers(-3, 4)))
    j = B.make_job(org, ref, stride, off, w, h, (0, 0, 0, 0), (0, 0))
    pred = np.zeros((h, w), np.int16)
    oracle_lib.vo_pred_qpel(C.byref(j), dx, dy, dq[0], dq[1], 0, B.ptr(pred), w)
    org = np.clip(pred.astype(np.int32) + np.rint(rng.normal(0, 3, pred.shape)).astype(np.int32), 0, 1023).astype(np.int16)
    return ref, np.ascontiguousarray(org), off, stride, (dx, dy), dq


@pytest.mark.ref
def test_fractional_branches_match_reference(oracle_lib, ref_lib):
    """Sub-pel planted motion so that every (half, quarter) branch of xExtDIFUpSamplingQ is exercised."""
    rng = np.random.default_rng(4242)
    seen = set()
    for rep in range(240):
        w, h = [(8, 8), (16, 16), (8, 16), (16, 8), (4, 8), (8, 4), (32, 16), (16, 32)][rep % 8]
        sr = 12
        ref, org, off, stride, _, _ = planted_frac(oracle_lib, rng, w, h, sr)
        imv, alt = ((0, 0), (0, 0), (0, 0), (1, 1))[rep % 4]
        pred = (int(rng.integers(-40, 40)), int(rng.integers(-40, 40)))
        j = B.make_job(org, ref, stride, off, w, h, (-sr, sr, -sr, sr), pred, imv, 0, 10, 1, alt, 1, 12.5)
        r0, r1, r2 = B.Result(), B.Result(), B.Result()
        ref_lib.ref_search(C.byref(j), C.byref(r0))
        oracle_lib.vo_search(C.byref(j), C.byref(r1), 1)
        oracle_lib.vo_search(C.byref(j), C.byref(r2), 0)
        assert r0.tuple() == r1.tuple()
        assert r0.tuple() == r2.tuple()
        seen.add((r0.halfX, r0.halfY, r0.qterX, r0.qterY))
    assert len({s[:2] for s in seen}) >= 8      # (nearly) all nine half-pel outcomes
    assert len(seen) >= 30                      # and a good share of the 81 (half, quarter) combinations


# ---- motion compensation: the oracle against the reference's own InterPrediction::xPredInterBlk ---------------------
@pytest.mark.ref
@pytest.mark.parametrize("comp", [0, 1])
@pytest.mark.parametrize("bi", [0, 1])
@pytest.mark.parametrize("alt", [0, 1])
def test_mc_block_matches_xPredInterBlk(oracle_lib, ref_lib, comp, bi, alt):
    from tests.helpers import mc_cases, oracle_mc
    W, H, M = 256, 128, 96
    cw, ch = (W, H) if comp == 0 else (W // 2, H // 2)
    rng = np.random.default_rng(20 + comp)
    padded = np.ascontiguousarray(np.pad(rng.integers(0, 1024, (ch, cw), dtype=np.int16), M, mode="edge"))
    blks = mc_cases(31 + comp * 2 + bi, comp, cw, ch, 250)
    ba = np.array(blks, dtype=np.int32)
    got = np.zeros(int((ba[:, 2] * ba[:, 3]).sum()), np.int16)
    rc = ref_lib.ref_mc_blocks(comp, B.ptr(padded), padded.shape[1], W, H, M, len(blks), C.c_void_p(ba.ctypes.data), bi, 10,
                               alt, B.ptr(got), None)
    assert rc == 0
    want = oracle_mc(oracle_lib, comp, padded, M, blks, bi, 10, alt)
    assert np.array_equal(got, want)


@pytest.mark.ref
def test_mc_block_8bit(oracle_lib, ref_lib):
    from tests.helpers import mc_cases, oracle_mc
    W, H, M = 128, 64, 96
    rng = np.random.default_rng(5)
    padded = np.ascontiguousarray(np.pad(rng.integers(0, 256, (H, W), dtype=np.int16), M, mode="edge"))
    blks = mc_cases(6, 0, W, H, 120, sizes=[4, 8, 16, 32])
    ba = np.array(blks, dtype=np.int32)
    for bi in (0, 1):
        got = np.zeros(int((ba[:, 2] * ba[:, 3]).sum()), np.int16)
        assert ref_lib.ref_mc_blocks(0, B.ptr(padded), padded.shape[1], W, H, M, len(blks), C.c_void_p(ba.ctypes.data), bi, 8, 0,
                                     B.ptr(got), None) == 0
        assert np.array_equal(got, oracle_mc(oracle_lib, 0, padded, M, blks, bi, 8, 0))


@pytest.mark.ref
def test_add_avg_and_remove_high_freq(oracle_lib, ref_lib):
    rng = np.random.default_rng(9)
    for bd in (8, 10):
        for w, h in [(4, 4), (8, 16), (12, 8), (64, 64), (2, 8)]:
            # 14-bit intermediates of two bi predictions
            s0 = rng.integers(-8192, 8192, (h, w), dtype=np.int16)
            s1 = rng.integers(-8192, 8192, (h, w), dtype=np.int16)
            a = np.zeros((h, w), np.int16)
            b = np.zeros((h, w), np.int16)
            ref_lib.ref_add_avg(B.ptr(s0), B.ptr(s1), B.ptr(a), w, h, bd)
            oracle_lib.vo_add_avg(B.ptr(s0), B.ptr(s1), B.ptr(b), w * h, bd)
            assert np.array_equal(a, b), (bd, w, h)
            if w % 4:
                continue   # removeHighFreq's unclipped branch exists for widths that are multiples of 4 only (Buffer.h:486-491)
            for clip in (0, 1):
                org = rng.integers(0, 1 << bd, (h, w), dtype=np.int16)
                pred = rng.integers(0, 1 << bd, (h, w), dtype=np.int16)
                a, b = org.copy(), org.copy()
                ref_lib.ref_remove_high_freq(B.ptr(a), w, B.ptr(pred), w, w, h, clip, bd)
                oracle_lib.vo_remove_high_freq(B.ptr(b), B.ptr(pred), w * h, clip, bd)
                assert np.array_equal(a, b), (bd, w, h, clip)


@pytest.mark.ref
def test_template_distortion_composition(oracle_lib, ref_lib):
    """xGetTemplateCost's distortion (InterSearch.cpp:3252-3266): getDistPart(DF_SAD) on xPredInterBlk's output — the
    reference's two functions chained against the oracle's two."""
    W, H, M = 128, 64, 96
    rng = np.random.default_rng(41)
    padded = np.ascontiguousarray(np.pad(rng.integers(0, 1024, (H, W), dtype=np.int16), M, mode="edge"))
    stride = padded.shape[1]
    for w, h in [(8, 8), (16, 8), (4, 16), (32, 32), (64, 16), (128, 64)]:
        org = rng.integers(0, 1024, (h, w), dtype=np.int16)
        for rep in range(6):
            mvx, mvy = (int(v) for v in rng.integers(-300, 300, 2))
            blk = np.array([[0, 0, w, h, mvx, mvy]], dtype=np.int32)
            rp = np.zeros((h, w), np.int16)
            assert ref_lib.ref_mc_blocks(0, B.ptr(padded), stride, W, H, M, 1, C.c_void_p(blk.ctypes.data), 0, 10, 0, B.ptr(rp), None) == 0
            op = np.zeros((h, w), np.int16)
            oracle_lib.vo_mc_block(0, B.ptr(padded, M * stride + M), stride, w, h, mvx, mvy, 0, 10, 0, B.ptr(op), w)
            assert ref_lib.ref_dist(B.ptr(org), w, B.ptr(rp), w, w, h, 10, 0, 0) == oracle_lib.vo_sad(B.ptr(org), w, B.ptr(op), w, w, h, 0)


@pytest.mark.ref
@pytest.mark.parametrize("imv", [1, 2])
@pytest.mark.parametrize("use_had", [0, 1])
def test_int_refine(oracle_lib, ref_lib, imv, use_had):
    """xPatternSearchIntRefine (InterSearch.cpp:4172-4282): the reference's own member against the restatement, over
    random AMVP states, all CU shapes, positions at the picture border (clipMv) and bi-pred weights."""
    from tests.helpers import MARGIN, clone_int_refine, int_refine_case, pad_plane
    rng = np.random.default_rng(100 + 10 * imv + use_had)
    pic_w, pic_h = 192, 128
    ref = rng.integers(0, 1024, (pic_h, pic_w), dtype=np.int16)
    refp = pad_plane(ref)
    stride = refp.shape[1]
    n = 0
    for w, h in [(a, b) for a in (4, 8, 16, 32, 64, 128) for b in (4, 8, 16, 32, 64, 128) if a * b > 16]:
        for rep in range(6):
            x = int(rng.integers(0, (pic_w - w) // 4 + 1)) * 4
            y = int(rng.integers(0, (pic_h - h) // 4 + 1)) * 4
            org = np.ascontiguousarray(rng.integers(0, 1024, (h, w), dtype=np.int16))
            if rep == 5:   # bi-pred pattern range
                org = (2 * org.astype(np.int32) - rng.integers(0, 1024, org.shape)).astype(np.int16)
            io = int_refine_case(rng, imv, x, y, w, h, pic_w, pic_h, max_pel=24 if rep else 150)
            j = B.make_job(org, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, (0, 0, 0, 0), (0, 0), 0, 0, 10,
                           use_had, 0, 0, 31.33 if rep % 2 else 57.91)
            a, b = clone_int_refine(io), clone_int_refine(io)
            # the clip keeps the block within CTU+8 samples of the picture: inside the padded plane for this size
            oracle_lib.vo_int_refine(C.byref(j), C.byref(a))
            ref_lib.ref_int_refine(C.byref(j), C.byref(b))
            assert a.tuple() == b.tuple(), (w, h, rep, a.tuple(), b.tuple())
            n += 1
    assert n > 100


@pytest.mark.ref
@pytest.mark.parametrize("extended,fast", [(0, 0), (1, 0), (0, 1)])
def test_tz_search(oracle_lib, ref_lib, extended, fast):
    """xTZSearch (InterSearch.cpp:3640-3974): the reference's own member against the restatement — FastSearch=1
    (diamond), FastSearch=3 (enhanced) and the fast re-search; real motion so that raster and star refinement run;
    all CU shapes, FEN row sub-sampling, history seeds with duplicates, picture-border positions."""
    from tests.helpers import MARGIN, pad_plane, tz_case
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(200 + 2 * extended + fast)
    pic_w, pic_h = 256, 192
    n = probes = 0
    for seed in range(3):
        cur, ref, _ = make_pair(50 + seed, pic_w, pic_h, max_global=20, max_local=30, n_rects=4, sigma=6.0)
        refp = pad_plane(ref)
        stride = refp.shape[1]
        for w, h in [(a, b) for a in (4, 8, 16, 32, 64, 128) for b in (4, 8, 16, 32, 64, 128) if a * b > 16]:
            for rep in range(3):
                x = int(rng.integers(0, (pic_w - w) // 4 + 1)) * 4
                y = int(rng.integers(0, (pic_h - h) // 4 + 1)) * 4
                if rep == 2:
                    x, y = [0, pic_w - w][int(rng.integers(0, 2))], [0, pic_h - h][int(rng.integers(0, 2))]
                sr = [64, 32, 96][rep]
                t = tz_case(rng, x, y, pic_w, pic_h, sr, extended, fast, first_stop=int(rep != 1),
                            max_pel=20 if rep < 2 else 160)
                pq = (int(rng.integers(-80, 81)), int(rng.integers(-80, 81)))
                j = B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, (0, 0, 0, 0), pq,
                               [0, 2, 0][rep], [0, 2, 2][rep], 10, 1, 0, 0, 31.33 if rep else 8.5, org_off=y * pic_w + x,
                               org_stride=pic_w)
                a = (C.c_int(), C.c_int(), C.c_uint64(), C.c_int())
                b = (C.c_int(), C.c_int(), C.c_uint64())
                oracle_lib.vo_tz_search(C.byref(j), C.byref(t), *[C.byref(v) for v in a])
                ref_lib.ref_tz_search(C.byref(j), C.byref(t), *[C.byref(v) for v in b])
                assert [v.value for v in a[:3]] == [v.value for v in b], (w, h, rep, seed)
                n += 1
                probes += a[3].value
    assert n > 300 and probes / n > 20


@pytest.mark.ref
@pytest.mark.parametrize("selective,fast", [(1, 0), (0, 1), (0, 0)])
def test_tz_search_selective(oracle_lib, ref_lib, selective, fast):
    """xTZSearchSelective (InterSearch.cpp:3979-4170, FastSearch=2) and xTZSearchHelp's staged SAD of subShiftMode 1
    (:340-391), which the selective method also switches on for the cached-MV re-search through xTZSearch (:3438-3441):
    the reference's own members against the restatement.  Large motion so that the full-search branch runs, small motion
    for the star refinement; all CU shapes (sub-shifts 1-4), border positions, sub-sampling modes 1 / 0 / 2."""
    from tests.helpers import MARGIN, pad_plane, tz_case
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(700 + 2 * selective + fast)
    pic_w, pic_h = 256, 192
    n = probes = far = 0
    for seed in range(3):
        cur, ref, _ = make_pair(80 + seed, pic_w, pic_h, max_global=[3, 14, 24][seed], max_local=30, n_rects=4, sigma=6.0)
        refp = pad_plane(ref)
        stride = refp.shape[1]
        for w, h in [(a, b) for a in (4, 8, 16, 32, 64, 128) for b in (4, 8, 16, 32, 64, 128) if a * b > 16]:
            for rep in range(3):
                x = int(rng.integers(0, (pic_w - w) // 4 + 1)) * 4
                y = int(rng.integers(0, (pic_h - h) // 4 + 1)) * 4
                if rep == 2:
                    x, y = [0, pic_w - w][int(rng.integers(0, 2))], [0, pic_h - h][int(rng.integers(0, 2))]
                sr = [64, 32, 96][rep]
                t = tz_case(rng, x, y, pic_w, pic_h, sr, 0, fast, first_stop=int(rep != 1), max_pel=20 if rep < 2 else 160)
                t.selective = selective
                pq = (int(rng.integers(-80, 81)), int(rng.integers(-80, 81)))
                mode = [1, 1, 2 * selective][rep] if (selective or fast) else 1
                j = B.make_job(cur, refp, stride, (MARGIN + y) * stride + MARGIN + x, w, h, (0, 0, 0, 0), pq,
                               [0, 2, 0][rep], mode, 10, 1, 0, 0, 31.33 if rep else 8.5, org_off=y * pic_w + x,
                               org_stride=pic_w)
                a = (C.c_int(), C.c_int(), C.c_uint64(), C.c_int())
                b = (C.c_int(), C.c_int(), C.c_uint64())
                oracle_lib.vo_tz_search(C.byref(j), C.byref(t), *[C.byref(v) for v in a])
                ref_lib.ref_tz_search(C.byref(j), C.byref(t), *[C.byref(v) for v in b])
                assert [v.value for v in a[:3]] == [v.value for v in b], (w, h, rep, seed)
                n += 1
                probes += a[3].value
                far += a[3].value > 2000
    assert n > 300 and probes / n > 20
    if selective:
        assert far > 10          # the full-search branch ran


@pytest.mark.ref
@pytest.mark.parametrize("bd", [10, 8])
def test_bcw_average_and_target(oracle_lib, ref_lib, bd):
    """The BCW forms of the bi-prediction helpers: AreaBuf::addWeightedAvg (Buffer.cpp:365-396) over 14-bit intermediates and
    AreaBuf::removeWeightHighFreq (Buffer.h:418-472; SIMD without clipping, scalar with), all five weights."""
    rng = np.random.default_rng(1300 + bd)
    for w, h in [(8, 8), (16, 4), (4, 8), (64, 32), (128, 128)]:
        n = w * h
        lo, hi = -8192, ((1 << bd) - 1 << (14 - bd)) - 8192                 # range of the bi = 1 predictions
        s0 = np.ascontiguousarray(rng.integers(lo, hi + 1, n).astype(np.int16))
        s1 = np.ascontiguousarray(rng.integers(lo, hi + 1, n).astype(np.int16))
        org = np.ascontiguousarray(rng.integers(0, 1 << bd, n).astype(np.int16))
        pred = np.ascontiguousarray(rng.integers(0, 1 << bd, n).astype(np.int16))
        for idx, wt in enumerate([-2, 3, 4, 5, 10]):
            a, b = np.zeros(n, np.int16), np.zeros(n, np.int16)
            oracle_lib.vo_add_weighted_avg(B.ptr(s0), B.ptr(s1), B.ptr(a), n, bd, idx)
            ref_lib.ref_add_weighted_avg(B.ptr(s0), B.ptr(s1), B.ptr(b), w, h, bd, idx)
            assert np.array_equal(a, b), (w, h, idx)
            for wsel in (wt, 8 - wt):                                       # the weight of either list
                if wsel == 4:
                    continue
                for clip in (0, 1):
                    a, b = org.copy(), org.copy()
                    oracle_lib.vo_remove_weight_high_freq(B.ptr(a), B.ptr(pred), n, clip, bd, wsel)
                    ref_lib.ref_remove_weight_high_freq(B.ptr(b), B.ptr(pred), w, h, clip, bd, wsel)
                    assert np.array_equal(a, b), (w, h, wsel, clip)


@pytest.mark.ref
@pytest.mark.parametrize("bd", [10, 8])
def test_dmvr_final_prediction(oracle_lib, ref_lib, bd):
    """The luma (and, for moved blocks, 4:2:0 chroma) prediction of both lists after DMVR (xPrefetch + xPad + xFinalPaddedMCForDMVR, InterPrediction.cpp:1664-1730,
    1845-1917): the reference's own members against the restatement (8-tap filter over the prefetched window with clamped
    coordinates), with the refinements the reference's own search found — integer moves of up to two samples, sub-sample
    steps, unmoved blocks — and MVs the clip moves at the picture border."""
    from tests.helpers import MARGIN, dmvr_cases, pad_plane
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(1200 + bd)
    pic_w, pic_h = 256, 192
    r0, r1, _ = make_pair(170, pic_w, pic_h, max_global=3, max_local=3, n_rects=4, sigma=4.0, bit_depth=bd)
    p0, p1 = pad_plane(r0), pad_plane(r1)
    stride = p0.shape[1]
    blk = dmvr_cases(rng, pic_w, pic_h, 500)
    res = np.zeros((len(blk), 4), np.int32)
    assert ref_lib.ref_dmvr_blocks(B.ptr(p0), B.ptr(p1), stride, pic_w, pic_h, MARGIN, len(blk), C.c_void_p(blk.ctypes.data), bd,
                                   C.c_void_p(res.ctypes.data)) == 0
    mvd = np.ascontiguousarray(res[:, :2])
    mvd[::5] = 0                                                   # unmoved blocks: no padding, luma still read from the window
    total = int((blk[:, 2] * blk[:, 3]).sum())
    want0, want1 = np.zeros(total, np.int16), np.zeros(total, np.int16)
    assert ref_lib.ref_dmvr_final_luma(B.ptr(p0), B.ptr(p1), stride, pic_w, pic_h, MARGIN, len(blk), C.c_void_p(blk.ctypes.data),
                                       C.c_void_p(mvd.ctypes.data), bd, B.ptr(want0), B.ptr(want1)) == 0
    off, pos = MARGIN * stride + MARGIN, 0
    for b, d in zip(blk, mvd):
        x, y, w, h, m0x, m0y, m1x, m1y = (int(v) for v in b)
        got = np.zeros(w * h, np.int16)
        for plane, want, (mx, my), sgn in ((p0, want0, (m0x, m0y), 1), (p1, want1, (m1x, m1y), -1)):
            oracle_lib.vo_dmvr_final_luma(B.ptr(plane, off), stride, x, y, w, h, mx, my, mx + sgn * int(d[0]), my + sgn * int(d[1]),
                                          pic_w, pic_h, 128, 128, bd, B.ptr(got))
            assert np.array_equal(got, want[pos:pos + w * h]), (b.tolist(), d.tolist(), sgn)
        pos += w * h
    assert (np.abs(mvd) >= 16).any(axis=1).sum() > 50 and (mvd % 16 != 0).any(axis=1).sum() > 50
    # 4:2:0 chroma of the moved blocks (an unmoved block is the plain motion compensation)
    c0, c1 = (np.ascontiguousarray(rng.integers(0, 1 << bd, (pic_h // 2, pic_w // 2), dtype=np.int16)) for _ in range(2))
    cp0, cp1 = pad_plane(c0, MARGIN // 2), pad_plane(c1, MARGIN // 2)
    cstride = cp0.shape[1]
    ctotal = total // 4
    wc0, wc1 = np.zeros(ctotal, np.int16), np.zeros(ctotal, np.int16)
    assert ref_lib.ref_dmvr_final(B.ptr(p0), B.ptr(p1), stride, B.ptr(cp0), B.ptr(cp1), cstride, pic_w, pic_h, MARGIN, len(blk),
                                  C.c_void_p(blk.ctypes.data), C.c_void_p(mvd.ctypes.data), bd, B.ptr(want0), B.ptr(want1), B.ptr(wc0),
                                  B.ptr(wc1)) == 0
    coff, pos, n_chroma = (MARGIN // 2) * cstride + MARGIN // 2, 0, 0
    for b, d in zip(blk, mvd):
        x, y, w, h, m0x, m0y, m1x, m1y = (int(v) for v in b)
        sz = (w // 2) * (h // 2)
        if d.any():
            got = np.zeros(sz, np.int16)
            for plane, want, (mx, my), sgn in ((cp0, wc0, (m0x, m0y), 1), (cp1, wc1, (m1x, m1y), -1)):
                oracle_lib.vo_dmvr_final_chroma(B.ptr(plane, coff), cstride, x, y, w, h, mx, my, mx + sgn * int(d[0]), my + sgn * int(d[1]),
                                                pic_w, pic_h, 128, 128, bd, B.ptr(got))
                assert np.array_equal(got, want[pos:pos + sz]), (b.tolist(), d.tolist(), sgn)
            n_chroma += 1
        pos += sz
    assert n_chroma > 200


@pytest.mark.ref
@pytest.mark.parametrize("num_refs,bd,qp,strength", [(4, 10, 32, 0.95), (2, 10, 27, 1.5), (3, 8, 37, 0.95), (1, 10, 22, 1.5)])
def test_mctf_bilateral(oracle_lib, ref_lib, num_refs, bd, qp, strength):
    """EncTemporalFilter::bilateralFilter (EncTemporalFilter.cpp:555-623): the reference's own member (applyMotion of every
    neighbour + the weighting, luma and 4:2:0 chroma) against vo_mctf_apply_motion + vo_mctf_bilateral — every sample equal.
    The weights are double-precision exp() values; the restatement sums them in the reference's order."""
    from tests.helpers import pad_plane
    from vtm_b200.synth import make_pair
    w, h = 208, 120
    rng = np.random.default_rng(1100 + num_refs)
    org, _, _ = make_pair(160, w, h, max_global=5, max_local=8, n_rects=3, sigma=5.0, bit_depth=bd)
    org = np.ascontiguousarray(org)
    org_c = np.ascontiguousarray(rng.integers(0, 1 << bd, (h // 2, w // 2), dtype=np.int16))
    offsets = [[-1], [-1, 1], [-2, -1, 1], [-2, -1, 1, 2]][num_refs - 1]
    refs_y, refs_c, fields = [], [], []
    orgp = pad_plane(org, 128)
    stride = orgp.shape[1]
    off = 128 * stride + 128
    for k in range(num_refs):
        r = np.clip(np.roll(org.astype(np.int32), (k + 1, -2 * k - 1), (0, 1)) + np.rint(rng.normal(0, 3 + 2 * k, org.shape)).astype(np.int32),
                    0, (1 << bd) - 1).astype(np.int16)
        c = np.clip(org_c.astype(np.int32) + np.rint(rng.normal(0, 6, org_c.shape)).astype(np.int32), 0, (1 << bd) - 1).astype(np.int16)
        refs_y.append(np.ascontiguousarray(r))
        refs_c.append(np.ascontiguousarray(c))
        mv = np.zeros((h // 4, w // 4, 3), np.int32)
        rp = pad_plane(refs_y[-1], 128)
        ref_lib.ref_mctf_me(B.ptr(orgp, off), stride, B.ptr(rp, off), stride, w, h, bd, C.c_void_p(mv.ctypes.data))
        fields.append(mv)
    mvs = np.ascontiguousarray(np.stack(fields))
    ptrs_y = (C.c_void_p * num_refs)(*[a.ctypes.data for a in refs_y])
    ptrs_c = (C.c_void_p * num_refs)(*[a.ctypes.data for a in refs_c])
    offs = (C.c_int * num_refs)(*offsets)
    want_y, want_c = np.zeros((h, w), np.int16), np.zeros((h // 2, w // 2), np.int16)
    ref_lib.ref_mctf_bilateral(B.ptr(org), B.ptr(org_c), ptrs_y, ptrs_c, num_refs, C.c_void_p(mvs.ctypes.data), offs, w, h, bd, qp,
                               strength, B.ptr(want_y), B.ptr(want_c))
    # restatement: motion-compensate every neighbour, then weight
    for comp, (o, planes, cw, ch, cs) in enumerate([(org, refs_y, w, h, 0), (org_c, refs_c, w // 2, h // 2, 1)]):
        corr = []
        for k in range(num_refs):
            pp = pad_plane(planes[k], 128 >> cs)
            m = 128 >> cs
            d = np.zeros((ch, cw), np.int16)
            oracle_lib.vo_mctf_apply_motion(B.ptr(pp, m * pp.shape[1] + m), pp.shape[1], cw, ch, cs, cs, C.c_void_p(mvs[k].ctypes.data),
                                            w // 4, bd, B.ptr(d), cw)
            corr.append(d)
        cptrs = (C.c_void_p * num_refs)(*[a.ctypes.data for a in corr])
        got = np.zeros((ch, cw), np.int16)
        oracle_lib.vo_mctf_bilateral(B.ptr(o), cw, cptrs, cw, offs, num_refs, cw, ch, comp, qp, strength, bd, B.ptr(got), cw)
        want = want_y if comp == 0 else want_c
        assert np.array_equal(got, want), (comp, np.argwhere(got != want)[:5])
        assert (got != o).mean() > 0.2          # the filter does change the picture
        # the table-driven form a device implementation would run (weights looked up by |ref - org|, IEEE double products and
        # sums in reference order, one division, round half away from zero) gives the same samples
        tables = []
        for index in (0, 1):
            tb = np.zeros(1 << bd, np.float64)
            oracle_lib.vo_mctf_bilateral_weights(comp, qp, strength, bd, num_refs, index, C.c_void_p(tb.ctypes.data))
            tables.append(tb)
        new_val, wsum = o.astype(np.float64), np.ones(o.shape, np.float64)
        for k in range(num_refs):
            wk = tables[min(1, abs(offsets[k]) - 1)][np.abs(corr[k].astype(np.int32) - o.astype(np.int32))]
            new_val = new_val + wk * corr[k].astype(np.float64)
            wsum = wsum + wk
        q = new_val / wsum
        t = np.trunc(q)
        lut = np.clip(t + (q - t >= 0.5), 0, (1 << bd) - 1).astype(np.int16)
        assert np.array_equal(lut, want), comp


@pytest.mark.ref
@pytest.mark.parametrize("imv", [0, 1, 2, 3])
def test_smvd_search(oracle_lib, ref_lib, imv):
    """xSymmetricMotionEstimation (InterSearch.cpp:4506-4518): the reference's own member against the restatement — every
    AMVR precision (the half-sample one with the alternative filter), SATD and SAD, clipped bi-prediction targets, all five
    BCW weights (removeWeightHighFreq, scalar with clipping and SIMD without), PU
    shapes 8x8 .. 64x64, positions at the picture border with MVs that the clip moves, start costs that let the diamond
    run several rounds."""
    from tests.helpers import MARGIN, pad_plane
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(1000 + imv)
    pic_w, pic_h = 256, 192
    n = moved = 0
    for seed in range(2):
        cur, ref0, _ = make_pair(140 + seed, pic_w, pic_h, max_global=4, max_local=6, n_rects=4, sigma=4.0)
        _, ref1, _ = make_pair(150 + seed, pic_w, pic_h, max_global=4, max_local=6, n_rects=4, sigma=4.0)
        cur = np.ascontiguousarray(cur)
        p0, p1 = pad_plane(ref0), pad_plane(ref1)
        stride = p0.shape[1]
        off = MARGIN * stride + MARGIN
        ios = []
        for w, h in [(8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (8, 32), (64, 16), (32, 8)]:
            for rep in range(4):
                x = int(rng.integers(0, (pic_w - w) // 4 + 1)) * 4
                y = int(rng.integers(0, (pic_h - h) // 4 + 1)) * 4
                span = 12 * 16
                if rep == 3:
                    x, y = [0, pic_w - w][int(rng.integers(0, 2))], [0, pic_h - h][int(rng.integers(0, 2))]
                    span = 170 * 16
                io = B.SmvdIo()
                io.x, io.y, io.w, io.h, io.picW, io.picH, io.maxCuW, io.maxCuH = x, y, w, h, pic_w, pic_h, 128, 128
                io.bd, io.imv = 10, imv
                unit = [4, 16, 64, 8][imv]                      # MVs and predictors on the AMVR grid, as the encoder has them
                io.curPredX, io.curPredY, io.tarPredX, io.tarPredY = (int(rng.integers(-span, span + 1)) // unit * unit for _ in range(4))
                dx, dy = (int(rng.integers(-6, 7)) * unit for _ in range(2))
                io.curMvX, io.curMvY = io.curPredX + dx, io.curPredY + dy
                io.tarMvX, io.tarMvY = io.tarPredX - dx, io.tarPredY - dy
                io.clipBiPred, io.useHad = int(rep == 1), int(rep != 2)
                io.bcwIdx = [2, 2, 2, 2, 0, 1, 3, 4][len(ios) % 8]           # BCW_DEFAULT and the four unequal weights
                io.lambda_ = [31.33, 8.5, 57.9, 31.33][rep]
                io.cost = [2 ** 40, w * h * 12, w * h * 5, 2 ** 40][rep]      # finite start costs: some rounds find nothing better
                ios.append(io)
        arr = (B.SmvdIo * len(ios))(*ios)
        want = (B.SmvdIo * len(ios))(*ios)
        assert ref_lib.ref_smvd_search(B.ptr(cur), pic_w, B.ptr(p0), B.ptr(p1), stride, MARGIN, len(ios), want) == 0
        for i in range(len(ios)):
            start = arr[i].tuple()
            oracle_lib.vo_smvd_search(B.ptr(cur, arr[i].y * pic_w + arr[i].x), pic_w, B.ptr(p0, off), B.ptr(p1, off), stride, C.byref(arr[i]))
            assert arr[i].tuple() == want[i].tuple(), (seed, i, arr[i].w, arr[i].h, start)
            n += 1
            moved += arr[i].tuple()[:2] != start[:2]
    assert n == 64 and moved > 30


@pytest.mark.ref
@pytest.mark.parametrize("bd", [10, 8])
def test_dmvr_blocks(oracle_lib, ref_lib, bd):
    """The DMVR search of a sub-block (InterPrediction.cpp:2098-2154): the reference's own xPrefetch, xinitMC (bilinear
    prediction), xDMVRCost, xBIPMVRefine and xDMVRSubPixelErrorSurface against the restatement — all phases of the
    bilinear filter, the early exit below w*h, integer and sub-sample outcomes, MVs clipped at the picture border."""
    from tests.helpers import MARGIN, dmvr_cases, pad_plane
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(900 + bd)
    pic_w, pic_h = 256, 192
    n = moved = sub = skipped = 0
    for seed in range(4):
        # the two references are the two sides of a pair: true motion between them, so that refinements are found;
        # seed 3: both lists read the same picture (exact matches: early exit, zero costs on the error surface)
        r0, r1, _ = make_pair(120 + seed, pic_w, pic_h, max_global=[1, 2, 6, 2][seed], max_local=3, n_rects=4,
                              sigma=[1.0, 4.0, 6.0, 3.0][seed], bit_depth=bd)
        if seed == 3:
            r1 = r0
        p0, p1 = pad_plane(r0), pad_plane(r1)
        stride = p0.shape[1]
        blk = dmvr_cases(rng, pic_w, pic_h, 400, same=seed == 3)
        want = np.zeros((len(blk), 4), np.int32)
        assert ref_lib.ref_dmvr_blocks(B.ptr(p0), B.ptr(p1), stride, pic_w, pic_h, MARGIN, len(blk), C.c_void_p(blk.ctypes.data), bd,
                                       C.c_void_p(want.ctypes.data)) == 0
        got = np.zeros(4, np.int32)
        off = MARGIN * stride + MARGIN
        for i, b in enumerate(blk):
            oracle_lib.vo_dmvr_block(B.ptr(p0, off), B.ptr(p1, off), stride, *[int(v) for v in b], pic_w, pic_h, 128, 128, bd,
                                     C.c_void_p(got.ctypes.data))
            assert got.tolist() == want[i].tolist(), (seed, i, b.tolist())
        n += len(blk)
        moved += int(((want[:, 0] != 0) | (want[:, 1] != 0)).sum())
        sub += int(((want[:, 0] % 16 != 0) | (want[:, 1] % 16 != 0)).sum())
        skipped += int((want[:, 3] == 0).sum())
    assert n == 1600 and moved > 100 and sub > 50 and skipped > 20


@pytest.mark.ref
@pytest.mark.parametrize("w,h,bd", [(208, 120, 10), (176, 144, 8), (96, 64, 10)])
def test_mctf_motion_estimation(oracle_lib, ref_lib, w, h, bd):
    """EncTemporalFilter::motionEstimation (EncTemporalFilter.cpp:448-466): the reference's own member against the
    restatement — the whole pyramid (1/4, 1/2, full 16x16, full 8x8 with 1/16-sample refinement), every MV and error."""
    from tests.helpers import pad_plane
    from vtm_b200.synth import make_pair
    cur, ref, _ = make_pair(300 + w, w, h, max_global=9, max_local=14, n_rects=3, sigma=5.0, bit_depth=bd)
    curp, refp = pad_plane(cur, 128), pad_plane(ref, 128)
    stride = curp.shape[1]
    off = 128 * stride + 128
    a = np.zeros((h // 4, w // 4, 3), np.int32)
    b = np.zeros_like(a)
    oracle_lib.vo_mctf_me(B.ptr(curp, off), stride, B.ptr(refp, off), stride, w, h, bd, C.c_void_p(a.ctypes.data))
    ref_lib.ref_mctf_me(B.ptr(curp, off), stride, B.ptr(refp, off), stride, w, h, bd, C.c_void_p(b.ctypes.data))
    assert np.array_equal(a, b), np.argwhere(a != b)[:5]
    n_blocks = int((a[:, :, 2] != np.iinfo(np.int32).max).sum())
    assert n_blocks == ((h - 1) // 8) * ((w - 1) // 8)          # blocks with blockX + 8 < width, blockY + 8 < height
    assert (a[:, :, :2][a[:, :, 2] != np.iinfo(np.int32).max] % 16 != 0).any()   # fractional MVs occur


@pytest.mark.ref
@pytest.mark.parametrize("w,h,bd", [(208, 120, 10), (96, 64, 8)])
def test_mctf_apply_motion(oracle_lib, ref_lib, w, h, bd):
    """EncTemporalFilter::applyMotion (EncTemporalFilter.cpp:470-552), luma and 4:2:0 chroma, with the vectors of the
    reference's own motion estimation plus random 1/16-sample vectors (every filter phase, negative vectors)."""
    from tests.helpers import pad_plane
    from vtm_b200.synth import make_pair
    rng = np.random.default_rng(500 + w)
    cur, ref, _ = make_pair(310 + w, w, h, max_global=9, max_local=14, n_rects=3, sigma=5.0, bit_depth=bd)
    chroma = np.ascontiguousarray(rng.integers(0, 1 << bd, (h // 2, w // 2), dtype=np.int16))
    curp, refp = pad_plane(cur, 128), pad_plane(ref, 128)
    stride = curp.shape[1]
    off = 128 * stride + 128
    mv = np.zeros((h // 4, w // 4, 3), np.int32)
    ref_lib.ref_mctf_me(B.ptr(curp, off), stride, B.ptr(refp, off), stride, w, h, bd, C.c_void_p(mv.ctypes.data))
    for variant in range(2):
        if variant:
            mv[:, :, 0] = rng.integers(-300, 301, mv.shape[:2])
            mv[:, :, 1] = rng.integers(-300, 301, mv.shape[:2])
        want_y, want_c = np.zeros((h, w), np.int16), np.zeros((h // 2, w // 2), np.int16)
        ref_lib.ref_mctf_apply_motion(B.ptr(np.ascontiguousarray(ref)), B.ptr(chroma), w, h, bd, C.c_void_p(mv.ctypes.data),
                                      B.ptr(want_y), B.ptr(want_c))
        got_y, got_c = np.zeros_like(want_y), np.zeros_like(want_c)
        chp = pad_plane(chroma, 64)
        oracle_lib.vo_mctf_apply_motion(B.ptr(refp, off), stride, w, h, 0, 0, C.c_void_p(mv.ctypes.data), w // 4, bd, B.ptr(got_y), w)
        oracle_lib.vo_mctf_apply_motion(B.ptr(chp, 64 * chp.shape[1] + 64), chp.shape[1], w // 2, h // 2, 1, 1,
                                        C.c_void_p(mv.ctypes.data), w // 4, bd, B.ptr(got_c), w // 2)
        assert np.array_equal(got_y, want_y) and np.array_equal(got_c, want_c), variant


@pytest.mark.ref
@pytest.mark.parametrize("six", [0, 1])
def test_affine_gradient_primitives(oracle_lib, ref_lib, six):
    """AffineGradientSearch's three dispatch-table entries (CommonLib/AffineGradientSearch.cpp:64-174, the reference's SIMD
    versions where it installs them): Sobel derivatives of a prediction with their replicated border, and the normal-equation
    sums of xEqualCoeffComputer (4- and 6-parameter model) — every derivative and every int64 coefficient equal."""
    rng = np.random.default_rng(1300 + six)
    for (w, h) in [(16, 16), (32, 16), (16, 64), (64, 64), (128, 32), (128, 128)]:
        stride = w + 8
        pred = rng.integers(0, 1024, (h, stride)).astype(np.int16)
        # the residual has the row stride of the derivatives, as in xAffineMotionEstimation (both `width`; the reference's SIMD
        # entry indexes it with the derivatives' index)
        res = rng.integers(-1023, 1024, (h, w)).astype(np.int16)
        got, want = [np.zeros((h, w), np.int32) for _ in range(2)], [np.zeros((h, w), np.int32) for _ in range(2)]
        for v in (0, 1):
            oracle_lib.vo_affine_sobel(v, B.ptr(pred), stride, C.c_void_p(got[v].ctypes.data), w, w, h)
            ref_lib.ref_affine_sobel(v, B.ptr(pred), stride, C.c_void_p(want[v].ctypes.data), w, w, h)
            assert np.array_equal(got[v], want[v]), (w, h, v)
        cg, cw = np.zeros((7, 7), np.int64), np.zeros((7, 7), np.int64)
        cg[1, 0] = cw[1, 0] = 12345                     # the entries accumulate
        oracle_lib.vo_affine_equal_coeff(B.ptr(res), w, C.c_void_p(got[0].ctypes.data), C.c_void_p(got[1].ctypes.data), w,
                                         C.c_void_p(cg.ctypes.data), w, h, six)
        ref_lib.ref_affine_equal_coeff(B.ptr(res), w, C.c_void_p(want[0].ctypes.data), C.c_void_p(want[1].ctypes.data), w,
                                       C.c_void_p(cw.ctypes.data), w, h, six)
        assert np.array_equal(cg, cw), (w, h)
        assert np.abs(cg).max() > 1 << 32
