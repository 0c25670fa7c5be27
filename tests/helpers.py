"""Shared test helpers: the CPU side of the parity checks (oracle-driven)."""
import ctypes as C

import numpy as np

from oracle import bindings as B

MARGIN = 192


def pad_plane(plane, margin=MARGIN):
    """Replicated border, what Picture::extendPicBorder produces."""
    return np.ascontiguousarray(np.pad(plane, margin, mode="edge"))


def oracle_window(L, pred_q, x, y, pic_w, pic_h, sr, ctu=128):
    l, r, t, b = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    L.vo_set_search_range(int(pred_q[0]) * 4, int(pred_q[1]) * 4, x, y, pic_w, pic_h, ctu, ctu, sr,
                          C.byref(l), C.byref(r), C.byref(t), C.byref(b))
    return l.value, r.value, t.value, b.value


def oracle_frame_search(L, cur, ref_padded, margin, sr, lam, pred_q=None, frac=1, use_had=1, levels=range(5),
                        only=None, literal=0, sub_shift_mode=0, bit_depth=10):
    """Every grid-aligned square CU (level-major, raster) through vo_search.  Returns list of result tuples
    (mvQx, mvQy, intX, intY, intSad, fracCost) indexed like the frame API; `only` = subset of CU indices."""
    h, w = cur.shape
    stride = ref_padded.shape[1]
    out = {}
    idx = 0
    for l in range(5):
        s = 8 << l
        for cy in range(h // s):
            for cx in range(w // s):
                if l in levels and (only is None or idx in only):
                    x, y = cx * s, cy * s
                    pq = (0, 0) if pred_q is None else (int(pred_q[idx][0]), int(pred_q[idx][1]))
                    win = oracle_window(L, pq, x, y, w, h, sr)
                    j = B.make_job(cur, ref_padded, stride, (margin + y) * stride + margin + x, s, s, win, pq, 0,
                                   sub_shift_mode, bit_depth, use_had, 0, frac, lam, org_off=y * w + x, org_stride=w)
                    r = B.Result()
                    L.vo_search(C.byref(j), C.byref(r), literal)
                    out[idx] = (4 * r.mvX + 2 * r.halfX + r.qterX, 4 * r.mvY + 2 * r.halfY + r.qterY, r.mvX, r.mvY,
                                int(r.intSad), int(r.fracCost) if frac else int(r.intSad) + int(
                                    L.vo_mv_cost(lam, L.vo_mv_bits(r.mvX, r.mvY, pq[0], pq[1], 2, 0))))
                idx += 1
    return out


def gpu_tuple(rec):
    return (int(rec["mvQx"]), int(rec["mvQy"]), int(rec["intX"]), int(rec["intY"]), int(rec["intSad"]),
            int(rec["fracCost"]))


# ---- motion compensation -------------------------------------------------------------------------------------------
def mc_cases(seed, comp, cw, ch, n, sizes=None, max_mv_pel=40):
    """n blocks (x, y, w, h, mvX, mvY) inside a cw x ch component plane; MVs in 1/16 luma sample covering every
    branch of xPredInterBlk (integer, horizontal-only, vertical-only, two-stage, half-pel)."""
    rng = np.random.default_rng(seed)
    if sizes is None:
        sizes = [4, 8, 16, 32, 64, 128] if comp == 0 else [2, 4, 8, 16, 32, 64]
    unit = 16 if comp == 0 else 32
    out = []
    for i in range(n):
        w = min(int(rng.choice(sizes)), cw)
        h = min(int(rng.choice(sizes)), ch)
        x = int(rng.integers(0, (cw - w) // w + 1)) * w
        y = int(rng.integers(0, (ch - h) // h + 1)) * h
        mvx = int(rng.integers(-max_mv_pel * unit, max_mv_pel * unit))
        mvy = int(rng.integers(-max_mv_pel * unit, max_mv_pel * unit))
        if i % 5 == 0:
            mvx &= ~(unit - 1)
        if i % 7 == 0:
            mvy &= ~(unit - 1)
        if i % 11 == 0:
            mvx = (mvx & ~(unit - 1)) | (unit // 2)
        if i % 13 == 0:
            mvy = (mvy & ~(unit - 1)) | (unit // 2)
        out.append((x, y, w, h, mvx, mvy))
    return out


def oracle_mc(L, comp, padded, margin, blocks, bi=0, bit_depth=10, alt=0):
    """vo_mc_block over the blocks; packed int16 like vtmme_mc_host."""
    stride = padded.shape[1]
    parts = []
    for (x, y, w, h, mvx, mvy) in blocks:
        d = np.zeros((h, w), np.int16)
        L.vo_mc_block(comp, B.ptr(padded, (margin + y) * stride + margin + x), stride, w, h, mvx, mvy, bi, bit_depth,
                      alt, B.ptr(d), w)
        parts.append(d.ravel())
    return np.concatenate(parts)


# ---- integer / 4-pel AMVR refinement ---------------------------------------------------------------------------------
def round_amvr(v, imv):
    """Mv::roundTransPrecInternal2Amvr (Mv.h:216-219) on one component."""
    rs = 4 if imv == 1 else 6
    off = 1 << (rs - 1)
    q = (v + off - 1) >> rs if v >= 0 else (v + off) >> rs
    return q << rs


def int_refine_case(rng, imv, x, y, w, h, pic_w, pic_h, max_pel=24):
    """Random state of one xPatternSearchIntRefine call: integer-pel MV, two AMVP candidates rounded to the AMVR
    precision, MVP index bits and incoming ruiBits."""
    mv = [int(rng.integers(-max_pel, max_pel + 1)) * 16 for _ in range(2)]
    cands = [[round_amvr(int(m + rng.integers(-80, 81)), imv) for m in mv] for _ in range(2)]
    if rng.integers(0, 4) == 0:
        cands[1] = list(cands[0])
    num = 2 if rng.integers(0, 5) else 1
    idx = int(rng.integers(0, num))
    idx_bits = [1, 1] if num == 2 else [0, 0]
    io = B.IntRefine()
    io.imv, io.mvX, io.mvY, io.numCand, io.mvpIdx = imv, mv[0], mv[1], num, idx
    for i in range(2):
        io.candX[i], io.candY[i] = cands[i]
        io.mvpIdxBits[i] = idx_bits[i]
    io.bits = int(rng.integers(3, 12)) + idx_bits[idx]
    io.fWeight = float(rng.choice([1.0, 0.5, 0.375, 0.625]))
    io.posX, io.posY, io.picW, io.picH, io.maxCuW, io.maxCuH = x, y, pic_w, pic_h, 128, 128
    return io


def clone_int_refine(io):
    c = B.IntRefine()
    C.memmove(C.byref(c), C.byref(io), C.sizeof(io))
    return c


# ---- TZ search -------------------------------------------------------------------------------------------------------
def tz_case(rng, x, y, pic_w, pic_h, search_range, extended=0, fast=0, first_stop=1, max_pel=20, n_seeds=None):
    """Random state of one xTZSearch call: start MV (the AMVP predictor, 1/16 sample), optional 2Nx2N integer MV,
    history MVs with duplicates."""
    t = B.TzParams()
    t.startX, t.startY = (int(rng.integers(-max_pel * 16, max_pel * 16 + 1)) for _ in range(2))
    t.hasInt2Nx2N = int(rng.integers(0, 2))
    t.int2Nx2NX, t.int2Nx2NY = (int(rng.integers(-max_pel, max_pel + 1)) for _ in range(2))
    t.nSeeds = int(rng.integers(0, 16)) if n_seeds is None else n_seeds
    for i in range(t.nSeeds):
        if i and rng.integers(0, 3) == 0:
            k = int(rng.integers(0, i))
            t.seedX[i], t.seedY[i] = t.seedX[k], t.seedY[k]
        else:
            t.seedX[i], t.seedY[i] = (int(rng.integers(-max_pel * 16, max_pel * 16 + 1)) for _ in range(2))
    t.searchRange, t.extended, t.fast, t.firstSearchStop = search_range, extended, fast, first_stop
    t.posX, t.posY, t.picW, t.picH, t.maxCuW, t.maxCuH = x, y, pic_w, pic_h, 128, 128
    return t


def oracle_frame_tz(L, cur, ref_padded, margin, sr, lam, pred_q=None, fast_search=1, first_stop=1, sub_shift_mode=0):
    """Every grid-aligned square CU through vo_tz_search (started at its predictor, no history seeds) + the fractional
    refinement; result tuples like oracle_frame_search."""
    h, w = cur.shape
    stride = ref_padded.shape[1]
    out = {}
    idx = 0
    for l in range(5):
        s = 8 << l
        for cy in range(h // s):
            for cx in range(w // s):
                x, y = cx * s, cy * s
                pq = (0, 0) if pred_q is None else (int(pred_q[idx][0]), int(pred_q[idx][1]))
                t = B.TzParams()
                t.startX, t.startY = pq[0] * 4, pq[1] * 4
                t.searchRange, t.extended, t.fast, t.firstSearchStop = sr, int(fast_search == 3), 0, first_stop
                t.selective = int(fast_search == 2)
                t.posX, t.posY, t.picW, t.picH, t.maxCuW, t.maxCuH = x, y, w, h, 128, 128
                j = B.make_job(cur, ref_padded, stride, (margin + y) * stride + margin + x, s, s, (0, 0, 0, 0), pq, 0,
                               sub_shift_mode, 10, 1, 0, 1, lam, org_off=y * w + x, org_stride=w)
                mx, my, sad = C.c_int(), C.c_int(), C.c_uint64()
                L.vo_tz_search(C.byref(j), C.byref(t), C.byref(mx), C.byref(my), C.byref(sad), None)
                hx, hy, qx, qy, cost = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
                L.vo_frac_direct(C.byref(j), mx.value, my.value, C.byref(hx), C.byref(hy), C.byref(qx), C.byref(qy),
                                 C.byref(cost))
                out[idx] = (4 * mx.value + 2 * hx.value + qx.value, 4 * my.value + 2 * hy.value + qy.value, mx.value,
                            my.value, int(sad.value), int(cost.value))
                idx += 1
    return out


def dmvr_cases(rng, pic_w, pic_h, n, max_pel=24, same=False):
    """n random DMVR sub-blocks {x, y, w, h, mvL0, mvL1} (int32 [n, 8]): sizes 8 / 16, merge MVs in 1/16 sample around a
    mirrored pair (DMVR's bi-prediction with equal POC distances) plus a small mismatch, some integer / half-sample MVs,
    some blocks at the picture border with MVs pointing far outside (the MV clip of xPrefetch / xinitMC)."""
    blk = np.zeros((n, 8), np.int32)
    for i in range(n):
        w, h = int(rng.choice([8, 16])), int(rng.choice([8, 16]))
        x = int(rng.integers(0, (pic_w - w) // 4 + 1)) * 4
        y = int(rng.integers(0, (pic_h - h) // 4 + 1)) * 4
        mx, my = (int(rng.integers(-max_pel * 16, max_pel * 16 + 1)) for _ in range(2))
        kind = i % 8
        if kind == 1:
            mx, my = mx & ~15, my & ~15
        elif kind == 2:
            mx = mx & ~15
        elif kind == 3:
            my = (my & ~15) | 8
        ex, ey = (int(rng.integers(-40, 41)) for _ in range(2))
        if kind == 7:
            x, y = [0, pic_w - w][int(rng.integers(0, 2))], [0, pic_h - h][int(rng.integers(0, 2))]
            mx, my = (int(rng.integers(-200 * 16, 200 * 16 + 1)) for _ in range(2))
        blk[i] = (x, y, w, h, mx, my, -mx + ex, -my + ey)
        if same:
            # both lists read the same picture: list 1 = list 0 displaced by 2*(ox, oy) samples makes offset (ox, oy) exact
            ox, oy = (int(rng.integers(-2, 3)) for _ in range(2))
            if kind in (0, 4):
                ox = oy = 0          # identical predictions: the early exit below w*h
            blk[i, 6:8] = (mx + 32 * ox, my + 32 * oy)
    return blk
