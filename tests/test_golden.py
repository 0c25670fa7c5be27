"""CPU: the oracle against the committed golden vectors (tests/golden/vtm_golden.npz, generated from the
compiled reference by tests/golden/make_golden.py).  Runs where /root/reference does not exist."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import bindings as B

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vtm_golden.npz"))


def iter_dist():
    off = 0
    for (w, h, kind), vals in zip(G["dist_meta"], G["dist_val"]):
        n = int(w) * int(h)
        org = np.ascontiguousarray(G["dist_org"][off:off + n].reshape(h, w))
        cur = np.ascontiguousarray(G["dist_cur"][off:off + n].reshape(h, w))
        off += n
        yield int(w), int(h), int(kind), org, cur, [int(v) for v in vals]


def iter_interp():
    off = 0
    for m in G["if_meta"]:
        comp, w, h, frac, vert, first, last, alt = (int(v) for v in m)
        want = G["if_out"][off:off + w * h].reshape(h, w)
        off += w * h
        yield comp, w, h, frac, vert, first, last, alt, want


def iter_search():
    off = 0
    for m, lam, res in zip(G["search_meta"], G["search_lambda"], G["search_res"]):
        w, h, x, y, l, r, t, b, pqx, pqy, imv, alt, ssm = (int(v) for v in m)
        org = np.ascontiguousarray(G["search_org"][off:off + w * h].reshape(h, w))
        off += w * h
        yield w, h, x, y, (l, r, t, b), (pqx, pqy), imv, alt, ssm, float(lam), org, tuple(int(v) for v in res)


def test_golden_was_made_with_simd_reference():
    assert int(G["simd_level"][0]) >= 1


def test_dist_golden(oracle_lib):
    n = 0
    for w, h, kind, org, cur, vals in iter_dist():
        assert oracle_lib.vo_sad(B.ptr(org), w, B.ptr(cur), w, w, h, 0) == vals[0]
        assert oracle_lib.vo_sad(B.ptr(org), w, B.ptr(cur), w, w, h, oracle_lib.vo_subshift(2, w, h)) == vals[1]
        assert oracle_lib.vo_satd(B.ptr(org), w, B.ptr(cur), w, w, h) == vals[2]
        n += 1
    assert n == 72


def test_mv_rate_golden(oracle_lib):
    for x, y, px, py, scale, imv, bits in G["mv_bits"]:
        assert oracle_lib.vo_mv_bits(int(x), int(y), int(px), int(py), int(scale), int(imv)) == int(bits)
    for lam, row in zip(G["mv_cost_lambda"], G["mv_cost"]):
        for b, c in enumerate(row):
            assert oracle_lib.vo_mv_cost(float(lam), b) == int(c)


def test_interp_golden(oracle_lib):
    src, mid = np.ascontiguousarray(G["if_src"]), np.ascontiguousarray(G["if_mid"])
    for comp, w, h, frac, vert, first, last, alt, want in iter_interp():
        s = src if first else mid
        got = np.zeros((h, w), np.int16)
        if vert:
            oracle_lib.vo_filter_ver(comp, B.ptr(s, 8 * 40 + 8), 40, B.ptr(got), w, w, h, frac, first, last, 10, alt)
        else:
            oracle_lib.vo_filter_hor(comp, B.ptr(s, 8 * 40 + 8), 40, B.ptr(got), w, w, h, frac, last, 10, alt)
        assert np.array_equal(got, want), (comp, w, h, frac, vert, first, last, alt)


@pytest.mark.parametrize("literal", [1, 0])
def test_search_golden(oracle_lib, literal):
    plane = np.ascontiguousarray(G["search_ref"])
    W = plane.shape[1]
    for w, h, x, y, win, pq, imv, alt, ssm, lam, org, want in iter_search():
        j = B.make_job(org, plane, W, y * W + x, w, h, win, pq, imv, ssm, 10, 1, alt, 1, lam)
        r = B.Result()
        oracle_lib.vo_search(C.byref(j), C.byref(r), literal)
        assert r.tuple() == want, (w, h, imv, ssm, literal)


# ---- motion compensation fixtures (tests/golden/mc_golden.npz, from the reference's xPredInterBlk) -------------------
GM = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mc_golden.npz"))


def iter_mc():
    """(comp, padded plane, margin, blocks, bi, alt, golden packed prediction)"""
    m = int(GM["dims"][2])
    for comp, name in ((0, "y"), (1, "c")):
        padded = np.ascontiguousarray(np.pad(GM["plane_" + name], m, mode="edge"))
        blks = [tuple(int(v) for v in b) for b in GM["blocks_" + name]]
        for bi in (0, 1):
            for alt in (0, 1):
                key = "pred_%s_bi%d_alt%d" % (name, bi, alt)
                if key in GM:
                    yield comp, padded, m, blks, bi, alt, GM[key]


def test_oracle_mc_golden():
    from tests.helpers import oracle_mc
    L = B.oracle()
    n = 0
    for comp, padded, m, blks, bi, alt, want in iter_mc():
        assert np.array_equal(oracle_mc(L, comp, padded, m, blks, bi, 10, alt), want), (comp, bi, alt)
        n += 1
    assert n == 6


def test_oracle_bipred_helpers_golden():
    L = B.oracle()
    s0, s1 = np.ascontiguousarray(GM["avg_s0"]), np.ascontiguousarray(GM["avg_s1"])
    d = np.zeros_like(s0)
    L.vo_add_avg(B.ptr(s0), B.ptr(s1), B.ptr(d), s0.size, 10)
    assert np.array_equal(d, GM["avg_out"])
    pred = np.ascontiguousarray(GM["hf_pred"])
    for clip in (0, 1):
        t = np.ascontiguousarray(GM["hf_org"]).copy()
        L.vo_remove_high_freq(B.ptr(t), B.ptr(pred), t.size, clip, 10)
        assert np.array_equal(t, GM["hf_out%d" % clip])


# ---- TZ search / AMVR refinement fixtures (tests/golden/amvr_tz_golden.npz, make_golden_amvr_tz.py) -------------------
GA = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "amvr_tz_golden.npz"))
GA_W, GA_H = 192, 128


# selective search / staged SAD fixtures (tests/golden/tz_selective_golden.npz, make_golden_tz_selective.py); same pictures
GS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tz_selective_golden.npz"))


def iter_tz(G=None):
    """(w, h, x, y, predQ, subShiftMode, oracle TzParams, lambda, want (mvX, mvY, sad, hx, hy, qx, qy, fracCost))"""
    from oracle import bindings as B
    G = GA if G is None else G
    for row, lam, res in zip(G["tz_int"], G["tz_lambda"], G["tz_res"]):
        row = [int(v) for v in row]
        w, h, x, y, pqx, pqy, ssm = row[:7]
        t = B.TzParams()
        t.startX, t.startY, t.hasInt2Nx2N, t.int2Nx2NX, t.int2Nx2NY, t.nSeeds = row[7:13]
        for i in range(16):
            t.seedX[i], t.seedY[i] = row[13 + i], row[29 + i]
        t.searchRange, t.extended, t.fast, t.firstSearchStop = row[45:49]
        t.selective = row[49] if len(row) > 49 else 0
        t.posX, t.posY, t.picW, t.picH, t.maxCuW, t.maxCuH = x, y, GA_W, GA_H, 128, 128
        yield w, h, x, y, (pqx, pqy), ssm, t, float(lam), tuple(int(v) for v in res)


def iter_amvr():
    """(w, h, x, y, predQ, window, imv, useHad, oracle IntRefine state (mv not set), lambda,
    want (intX, intY, intSad, mvX, mvY, mvpIdx, bits, cost))"""
    from oracle import bindings as B
    for row, f, res in zip(GA["ir_int"], GA["ir_f"], GA["ir_res"]):
        row = [int(v) for v in row]
        w, h, x, y, pqx, pqy = row[:6]
        io = B.IntRefine()
        io.imv, io.numCand = row[10], row[12]
        io.candX[0], io.candY[0], io.candX[1], io.candY[1] = row[13:17]
        io.mvpIdx, io.mvpIdxBits[0], io.mvpIdxBits[1], io.bits = row[17:21]
        io.fWeight = float(f[1])
        io.posX, io.posY, io.picW, io.picH, io.maxCuW, io.maxCuH = x, y, GA_W, GA_H, 128, 128
        yield w, h, x, y, (pqx, pqy), tuple(row[6:10]), row[10], row[11], io, float(f[0]), tuple(int(v) for v in res)


def _ga_planes():
    from tests.helpers import MARGIN, pad_plane
    cur = np.ascontiguousarray(GA["cur"])
    refp = pad_plane(np.ascontiguousarray(GA["ref"]))
    return cur, refp, MARGIN


@pytest.mark.parametrize("which", ["tz", "selective"])
def test_tz_golden(oracle_lib, which):
    """oracle xTZSearch + fractional refinement == the reference's (60 cases: FastSearch=1, enhanced, fast re-search);
    oracle xTZSearchSelective / staged SAD + refinement == the reference's (60 cases: FastSearch=2)"""
    from oracle import bindings as B
    cur, refp, m = _ga_planes()
    stride = refp.shape[1]
    n = 0
    for w, h, x, y, pq, ssm, t, lam, want in iter_tz(GA if which == "tz" else GS):
        j = B.make_job(cur, refp, stride, (m + y) * stride + m + x, w, h, (0, 0, 0, 0), pq, 0, ssm, 10, 1, 0, 1, lam,
                       org_off=y * GA_W + x, org_stride=GA_W)
        mx, my, sad = C.c_int(), C.c_int(), C.c_uint64()
        oracle_lib.vo_tz_search(C.byref(j), C.byref(t), C.byref(mx), C.byref(my), C.byref(sad), None)
        hx, hy, qx, qy, cost = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
        oracle_lib.vo_frac_direct(C.byref(j), mx.value, my.value, C.byref(hx), C.byref(hy), C.byref(qx), C.byref(qy), C.byref(cost))
        assert (mx.value, my.value, sad.value, hx.value, hy.value, qx.value, qy.value, cost.value) == want, (w, h, x, y)
        n += 1
    assert n == 60


def test_amvr_golden(oracle_lib):
    """oracle xPatternSearch + xPatternSearchIntRefine == the reference's (40 cases: FPEL / 4PEL, SATD / SAD)"""
    from oracle import bindings as B
    cur, refp, m = _ga_planes()
    stride = refp.shape[1]
    n = 0
    for w, h, x, y, pq, win, imv, use_had, io, lam, want in iter_amvr():
        j = B.make_job(cur, refp, stride, (m + y) * stride + m + x, w, h, win, pq, imv << 1, 0, 10, use_had, 0, 0, lam,
                       org_off=y * GA_W + x, org_stride=GA_W)
        r = B.Result()
        oracle_lib.vo_search(C.byref(j), C.byref(r), 0)
        io.mvX, io.mvY = r.mvX * 16, r.mvY * 16
        oracle_lib.vo_int_refine(C.byref(j), C.byref(io))
        assert (r.mvX, r.mvY, r.intSad) + io.tuple() == want, (w, h, x, y)
        n += 1
    assert n == 40


def test_mctf_golden(oracle_lib):
    """oracle EncTemporalFilter::motionEstimation == the reference's (tests/golden/mctf_golden.npz)"""
    from oracle import bindings as B
    from tests.helpers import pad_plane
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mctf_golden.npz"))
    cur, ref = np.ascontiguousarray(g["cur"]), np.ascontiguousarray(g["ref"])
    h, w = cur.shape
    curp, refp = pad_plane(cur, 128), pad_plane(ref, 128)
    stride = curp.shape[1]
    off = 128 * stride + 128
    out = np.zeros((h // 4, w // 4, 3), np.int32)
    oracle_lib.vo_mctf_me(B.ptr(curp, off), stride, B.ptr(refp, off), stride, w, h, 10, C.c_void_p(out.ctypes.data))
    assert np.array_equal(out, g["mv"])


# ---- DMVR fixtures (tests/golden/dmvr_golden.npz, make_golden_dmvr.py); pictures of amvr_tz_golden.npz -----------------
GD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dmvr_golden.npz"))


def iter_dmvr():
    """(padded list-0 plane, padded list-1 plane, margin, blocks int32 [n, 8], want int32 [n, 4])"""
    from tests.helpers import MARGIN, pad_plane
    p0, p1 = pad_plane(np.ascontiguousarray(GA["ref"])), pad_plane(np.ascontiguousarray(GA["cur"]))
    yield p0, p1, MARGIN, GD["pair_blk"], GD["pair_res"]
    yield p0, p0, MARGIN, GD["same_blk"], GD["same_res"]


def test_dmvr_golden(oracle_lib):
    """oracle DMVR sub-block search == the reference's own members (320 cases)"""
    n = 0
    for p0, p1, m, blk, want in iter_dmvr():
        stride, off = p0.shape[1], m * p0.shape[1] + m
        got = np.zeros(4, np.int32)
        for b, w in zip(blk, want):
            oracle_lib.vo_dmvr_block(B.ptr(p0, off), B.ptr(p1, off), stride, *[int(v) for v in b], GA_W, GA_H, 128, 128, 10,
                                     C.c_void_p(got.ctypes.data))
            assert got.tolist() == w.tolist(), b.tolist()
            n += 1
    assert n == 320
