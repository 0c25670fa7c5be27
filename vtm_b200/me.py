"""Host-side mirror of the reference's motion-search interface over the C ABI (include/vtmme.h).

Names follow the reference: a *picture* is a luma plane with border (Picture::getRecoBuf), a *job* is one
xMotionEstimation call's integer search + fractional refinement (IntTZSearchStruct / DistParam fields),
the frame search covers the quad-tree of square CUs of every CTU.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from .lib import (CAffineBlock, CAmvr, CCandJob, CDmvrBlock, CDmvrResult, CSmvd, CSmvdResult, CTz, CFrameParams, CJob, CMcBlock, CResult, ERR_NAMES, VtmmeError,
                  load_library)

# vtmme_cu_result
CU_RESULT_DTYPE = np.dtype([("mvQx", "<i2"), ("mvQy", "<i2"), ("intX", "<i2"), ("intY", "<i2"),
                            ("intSad", "<u4"), ("fracCost", "<u4")])


@dataclass
class FrameParams:
    searchRange: int = 64
    bitDepth: int = 10
    ctuSize: int = 128
    imvShift: int = 0
    useHad: int = 1
    fracMode: int = 1
    predSpread: int = 0
    lambdaMotion: float = 31.33
    subShiftMode: int = 0
    fastSearch: int = 0       # 0 full search, 1 TZ search, 3 enhanced TZ search (VTM's --FastSearch)
    tzFirstSearchStop: int = 1

    def c(self):
        return CFrameParams(self.searchRange, self.bitDepth, self.ctuSize, self.imvShift, self.useHad, self.fracMode,
                            self.predSpread, self.subShiftMode, self.lambdaMotion, self.fastSearch,
                            self.tzFirstSearchStop)


@dataclass
class Job:
    """One xMotionEstimation search: fields of IntTZSearchStruct / DistParam (EncoderLib/InterSearch.h:337-352)."""
    curPic: int
    refPic: int
    x: int
    y: int
    w: int
    h: int
    sr: tuple                 # (left, right, top, bottom), integer pel, inclusive
    predQ: tuple              # quarter-pel predictor
    imvShift: int = 0
    subShift: int = 0
    bitDepth: int = 10
    useHad: int = 1
    useAltHpel: int = 0
    fracMode: int = 1
    lambdaMotion: float = 31.33
    org: np.ndarray = None    # optional int16 pattern override (bi-pred)
    amvr: "Amvr" = None       # fracMode 2: state of the xPatternSearchIntRefine call that follows the search
    tz: "TzSearch" = None     # integer search = xTZSearch (FastSearch=1/3) instead of the full search; `sr` is ignored


@dataclass
class TzSearch:
    """vtmme_tz: what xTZSearch (EncoderLib/InterSearch.cpp:3640-3974) or xTZSearchSelective (:3979-4170) receives;
    MVs in 1/16 sample."""
    start: tuple              # rcMv on entry
    searchRange: int
    picW: int
    picH: int
    seeds: tuple = ()         # history MVs, newest first
    int2Nx2N: tuple = None    # integer-pel MV or None
    extended: int = 0
    fast: int = 0
    firstSearchStop: int = 1
    maxCu: int = 128
    selective: int = 0        # xTZSearchSelective (:3979-4170, FastSearch=2) instead of xTZSearch
    stagedSad: int = 0        # subShiftMode 1: xTZSearchHelp's staged SAD (:340-391); Job.subShift = the mode-1 value

    def c(self):
        t = CTz()
        t.startX, t.startY = self.start
        t.hasInt2Nx2N = 0 if self.int2Nx2N is None else 1
        if self.int2Nx2N is not None:
            t.int2Nx2NX, t.int2Nx2NY = self.int2Nx2N
        t.nSeeds = len(self.seeds)
        for i, (x, y) in enumerate(self.seeds):
            t.seedX[i], t.seedY[i] = x, y
        t.searchRange, t.extended, t.fast, t.firstSearchStop = self.searchRange, self.extended, self.fast, self.firstSearchStop
        t.picW, t.picH, t.maxCu = self.picW, self.picH, self.maxCu
        t.selective, t.stagedSad = self.selective, self.stagedSad
        return t


@dataclass
class Amvr:
    """vtmme_amvr: what xPatternSearchIntRefine (EncoderLib/InterSearch.cpp:4172-4282) receives; MVs in 1/16 sample."""
    imv: int                  # 1 IMV_FPEL, 2 IMV_4PEL
    cands: tuple              # ((x0, y0), (x1, y1)) amvpInfo.mvCand
    numCand: int
    mvpIdx: int
    mvpIdxBits: tuple
    bits: int
    picW: int
    picH: int
    fWeight: float = 1.0
    maxCuW: int = 128
    maxCuH: int = 128

    def c(self):
        return CAmvr(self.imv, self.numCand, (C.c_int32 * 2)(self.cands[0][0], self.cands[1][0]),
                     (C.c_int32 * 2)(self.cands[0][1], self.cands[1][1]), self.mvpIdx,
                     (C.c_uint32 * 2)(*self.mvpIdxBits), self.bits, self.picW, self.picH, self.maxCuW, self.maxCuH,
                     self.fWeight)


def frame_cu_layout(width, height):
    """CU order of the frame search: (total, level offsets[6]); level l = CUs of size 8<<l in raster order."""
    off, acc = [], 0
    for l in range(5):
        s = 8 << l
        off.append(acc)
        acc += (width // s) * (height // s)
    off.append(acc)
    return acc, off


class MotionSearch:
    """One libvtmme context on one GPU."""

    def __init__(self, device=0):
        self.L = load_library()
        self.ctx = C.c_void_p()
        rc = self.L.vtmme_create(int(device), C.byref(self.ctx))
        if rc != 0:
            raise VtmmeError("vtmme_create(device=%d) failed: %s (no CUDA device? there is no CPU fallback)"
                             % (device, ERR_NAMES.get(rc, rc)))
        self.device = device

    def close(self):
        if self.ctx:
            self.L.vtmme_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise VtmmeError("%s failed: %s: %s" % (what, ERR_NAMES.get(rc, rc),
                                                     self.L.vtmme_last_error(self.ctx).decode()))

    # ---- stream / sync -------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._check(self.L.vtmme_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr)), "vtmme_set_stream")

    def synchronize(self):
        self._check(self.L.vtmme_synchronize(self.ctx), "vtmme_synchronize")

    def set_profiling(self, on=True):
        self._check(self.L.vtmme_set_profiling(self.ctx, 1 if on else 0), "vtmme_set_profiling")

    def frame_kernel_ms(self):
        """(me_tree_sad, me_tree_upper, me_frac_frame) milliseconds of the most recent profiled frame search."""
        ms = (C.c_float * 3)()
        self._check(self.L.vtmme_frame_kernel_ms(self.ctx, ms), "vtmme_frame_kernel_ms")
        return tuple(float(v) for v in ms)

    @property
    def launches(self):
        return int(self.L.vtmme_launch_count(self.ctx))

    # ---- pictures ------------------------------------------------------------------------------------------
    def upload_picture(self, pic_id, plane, margin=0):
        """plane: int16 array [height + 2*margin, width + 2*margin] (margin = valid border it already carries)."""
        assert plane.dtype == np.int16 and plane.ndim == 2 and plane.flags["C_CONTIGUOUS"]
        h, w = plane.shape[0] - 2 * margin, plane.shape[1] - 2 * margin
        origin = plane.ctypes.data + 2 * (margin * plane.shape[1] + margin)
        self._check(self.L.vtmme_upload_picture(self.ctx, pic_id, C.c_void_p(origin), plane.shape[1], w, h, margin,
                                                1 if margin else 0), "vtmme_upload_picture")

    def upload_picture_async(self, pic_id, host_ptr, stride, width, height, margin=0):
        """Pipelined upload from PAGE-LOCKED host memory (address of sample (0,0)); overlaps with searches on other
        pictures.  Keep the buffer alive until synchronize()."""
        self._check(self.L.vtmme_upload_picture_async(self.ctx, pic_id, C.c_void_p(host_ptr), stride, width, height, margin,
                                                      1 if margin else 0), "vtmme_upload_picture_async")

    def upload_picture_device(self, pic_id, dptr, stride, width, height, margin=0):
        """dptr: device address of sample (0,0) of an int16 plane (e.g. a torch tensor's data_ptr())."""
        self._check(self.L.vtmme_upload_picture_device(self.ctx, pic_id, C.c_void_p(dptr), stride, width, height,
                                                       margin, 1 if margin else 0), "vtmme_upload_picture_device")

    def release_picture(self, pic_id):
        self._check(self.L.vtmme_release_picture(self.ctx, pic_id), "vtmme_release_picture")

    # ---- per-call jobs ---------------------------------------------------------------------------------------
    def search(self, jobs):
        n = len(jobs)
        cj = (CJob * n)()
        keep = []
        for i, j in enumerate(jobs):
            org_ptr, org_stride = None, 0
            if j.org is not None:
                o = np.ascontiguousarray(j.org, dtype=np.int16)
                keep.append(o)
                org_ptr, org_stride = o.ctypes.data, o.shape[1]
            amvr = None
            if j.amvr is not None:
                amvr = j.amvr.c()
                keep.append(amvr)
                amvr = C.pointer(amvr)
            tz = None
            if j.tz is not None:
                tz = j.tz.c()
                keep.append(tz)
                tz = C.pointer(tz)
            cj[i] = CJob(j.curPic, j.refPic, j.x, j.y, j.w, j.h, org_ptr, org_stride, j.sr[0], j.sr[1], j.sr[2],
                         j.sr[3], j.predQ[0], j.predQ[1], j.imvShift, j.subShift, j.bitDepth, j.useHad, j.useAltHpel,
                         j.fracMode, j.lambdaMotion, amvr, tz)
        res = (CResult * n)()
        self._check(self.L.vtmme_search(self.ctx, cj, n, res), "vtmme_search")
        return [r.tuple() + (r.amvr_tuple() if j.fracMode == 2 else ()) for r, j in zip(res, jobs)]

    # ---- batched frame search --------------------------------------------------------------------------------
    def search_frames(self, cur_ids, ref_ids, params, pred_q=None, out=None):
        """Host buffers.  pred_q: int16 [nPairs, nCU, 2] or None.  Returns a structured array [nPairs, nCU]; `out`
        (optional) is a caller-owned array of that shape to receive the results — page-locked memory makes the
        device-to-host copy a single DMA instead of a staged one."""
        n = len(cur_ids)
        cur = (C.c_int32 * n)(*cur_ids)
        ref = (C.c_int32 * n)(*ref_ids)
        prm = params.c()
        ncu = self._ncu
        if out is None:
            out = np.zeros((n, ncu), dtype=CU_RESULT_DTYPE)
        assert out.dtype == CU_RESULT_DTYPE and out.shape == (n, ncu) and out.flags["C_CONTIGUOUS"]
        pp = None
        if pred_q is not None:
            pred_q = np.ascontiguousarray(pred_q, dtype=np.int16)
            assert pred_q.shape == (n, ncu, 2)
            pp = C.c_void_p(pred_q.ctypes.data)
        self._check(self.L.vtmme_search_frames(self.ctx, n, cur, ref, C.byref(prm), pp, C.c_void_p(out.ctypes.data)),
                    "vtmme_search_frames")
        return out

    def search_frames_device(self, cur_ids, ref_ids, params, d_pred_q, d_results):
        """Device buffers (addresses), asynchronous on the context stream."""
        n = len(cur_ids)
        cur = (C.c_int32 * n)(*cur_ids)
        ref = (C.c_int32 * n)(*ref_ids)
        prm = params.c()
        self._check(self.L.vtmme_search_frames_device(self.ctx, n, cur, ref, C.byref(prm),
                                                      C.c_void_p(d_pred_q) if d_pred_q else None,
                                                      C.c_void_p(d_results)), "vtmme_search_frames_device")

    def set_frame_size(self, width, height):
        self._ncu, self._off = frame_cu_layout(width, height)
        lo = (C.c_int32 * 6)()
        n = self.L.vtmme_frame_cu_count(width, height, lo)
        assert n == self._ncu and list(lo) == self._off
        return self._ncu

    # ---- table-level ---------------------------------------------------------------------------------------
    def dist_host(self, kind, org, cur, sub_shift=0):
        """One block pair in host memory (the DistParam-hook flavour).  org, cur: int16 2-D arrays (views ok)."""
        assert org.dtype == np.int16 and cur.dtype == np.int16
        h, w = org.shape
        out = C.c_uint64()
        self._check(self.L.vtmme_dist_host(self.ctx, kind, C.c_void_p(org.ctypes.data), org.strides[0] // 2,
                                           C.c_void_p(cur.ctypes.data), cur.strides[0] // 2, w, h, sub_shift,
                                           C.byref(out)), "vtmme_dist_host")
        return out.value

    def dist_batch(self, kind, d_org, org_stride, org_blk, d_cur, cur_stride, cur_blk, w, h, sub_shift, n, d_out):
        self._check(self.L.vtmme_dist_batch(self.ctx, kind, C.c_void_p(d_org), org_stride, org_blk, C.c_void_p(d_cur),
                                            cur_stride, cur_blk, w, h, sub_shift, n, C.c_void_p(d_out)),
                    "vtmme_dist_batch")

    def interp_host(self, comp, vertical, src, src_off, src_stride, w, h, frac, is_first, is_last, bit_depth=10,
                    use_alt_hpel=0):
        """src: int16 array, src_off: element offset of the first output position.  Returns int16 [h, w]."""
        dst = np.zeros((h, w), np.int16)
        self._check(self.L.vtmme_interp_host(self.ctx, comp, vertical, C.c_void_p(src.ctypes.data + 2 * src_off),
                                             src_stride, C.c_void_p(dst.ctypes.data), w, w, h, frac, is_first, is_last,
                                             bit_depth, use_alt_hpel), "vtmme_interp_host")
        return dst

    def filter_host(self, taps, vertical, is_first, is_last, copy, src, src_off, src_stride, w, h, coeff, bit_depth=10):
        """Function-pointer flavour (explicit taps): what m_filterHor/m_filterVer/m_filterCopy entries receive."""
        dst = np.zeros((h, w), np.int16)
        c = np.ascontiguousarray(coeff, dtype=np.int16) if coeff is not None else None
        self._check(self.L.vtmme_filter_host(self.ctx, taps, vertical, is_first, is_last, copy,
                                             C.c_void_p(src.ctypes.data + 2 * src_off), src_stride,
                                             C.c_void_p(dst.ctypes.data), w, w, h,
                                             C.c_void_p(c.ctypes.data) if c is not None else None, bit_depth),
                    "vtmme_filter_host")
        return dst

    def interp_batch(self, comp, vertical, d_src, src_stride, src_blk, d_dst, dst_stride, dst_blk, w, h, frac, is_first,
                     is_last, bit_depth, use_alt_hpel, n):
        self._check(self.L.vtmme_interp_batch(self.ctx, comp, vertical, C.c_void_p(d_src), src_stride, src_blk,
                                              C.c_void_p(d_dst), dst_stride, dst_blk, w, h, frac, is_first, is_last,
                                              bit_depth, use_alt_hpel, n), "vtmme_interp_batch")

    # ---- motion compensation -------------------------------------------------------------------------------
    @staticmethod
    def _mc_blocks(blocks):
        """blocks: iterable of (refPic, x, y, w, h, mvX, mvY); mv in 1/16 luma sample."""
        arr = (CMcBlock * len(blocks))()
        total = 0
        for i, b in enumerate(blocks):
            arr[i] = CMcBlock(int(b[0]), int(b[1]), int(b[2]), int(b[3]), int(b[4]), int(b[5]), int(b[6]), 0)
            total += int(b[3]) * int(b[4])
        return arr, total

    def mc_host(self, comp, blocks, bi=0, bit_depth=10, use_alt_hpel=0):
        """xPredInterBlk for every block; returns the packed int16 predictions (block i at the sum of earlier w*h)."""
        arr, total = self._mc_blocks(blocks)
        dst = np.zeros(total, np.int16)
        self._check(self.L.vtmme_mc_host(self.ctx, comp, bi, bit_depth, use_alt_hpel, len(blocks), arr,
                                         C.c_void_p(dst.ctypes.data)), "vtmme_mc_host")
        return dst

    def mc_batch(self, comp, blocks, d_dst, bi=0, bit_depth=10, use_alt_hpel=0):
        """Same, into device memory at d_dst, asynchronous on the context stream."""
        arr, _ = self._mc_blocks(blocks) if not isinstance(blocks, C.Array) else (blocks, 0)
        self._check(self.L.vtmme_mc_batch(self.ctx, comp, bi, bit_depth, use_alt_hpel, len(arr), arr, C.c_void_p(d_dst)),
                    "vtmme_mc_batch")

    def add_avg(self, d_src0, d_src1, d_dst, count, bit_depth=10):
        self._check(self.L.vtmme_add_avg(self.ctx, C.c_void_p(d_src0), C.c_void_p(d_src1), C.c_void_p(d_dst), count,
                                         bit_depth), "vtmme_add_avg")

    def remove_high_freq(self, d_org, d_pred, count, clip=0, bit_depth=10):
        self._check(self.L.vtmme_remove_high_freq(self.ctx, C.c_void_p(d_org), C.c_void_p(d_pred), count, clip, bit_depth),
                    "vtmme_remove_high_freq")

    def add_weighted_avg(self, d_src0, d_src1, d_dst, count, bcw_idx, bit_depth=10):
        self._check(self.L.vtmme_add_weighted_avg(self.ctx, C.c_void_p(d_src0), C.c_void_p(d_src1), C.c_void_p(d_dst), count,
                                                  bit_depth, bcw_idx), "vtmme_add_weighted_avg")

    def remove_weight_high_freq(self, d_org, d_pred, count, bcw_weight, clip=0, bit_depth=10):
        self._check(self.L.vtmme_remove_weight_high_freq(self.ctx, C.c_void_p(d_org), C.c_void_p(d_pred), count, clip, bit_depth,
                                                         bcw_weight), "vtmme_remove_weight_high_freq")

    # ---- GOP-based temporal filter ---------------------------------------------------------------------------
    def mctf_me(self, org_ids, ref_ids, width, height, bit_depth=10):
        """EncTemporalFilter::motionEstimation for every (original, reference) pair of uploaded pictures.
        Returns int32 [nPairs, height/4, width/4, 3] = {x, y, error}, x / y in 1/16 sample."""
        n = len(org_ids)
        out = np.zeros((n, height // 4, width // 4, 3), np.int32)
        self._check(self.L.vtmme_mctf_me(self.ctx, n, (C.c_int32 * n)(*org_ids), (C.c_int32 * n)(*ref_ids), bit_depth,
                                         C.c_void_p(out.ctypes.data)), "vtmme_mctf_me")
        return out

    def mctf_apply_motion(self, src_id, comp_w, comp_h, mv, csx=0, csy=0, bit_depth=10):
        """EncTemporalFilter::applyMotion for one component plane (uploaded as src_id) with the luma vector field
        `mv` (int32 [rows, cols, 3] as returned by mctf_me).  Returns int16 [comp_h, comp_w]."""
        mv = np.ascontiguousarray(mv, dtype=np.int32)
        out = np.zeros((comp_h, comp_w), np.int16)
        self._check(self.L.vtmme_mctf_apply_motion(self.ctx, src_id, csx, csy, C.c_void_p(mv.ctypes.data), mv.shape[1],
                                                   mv.shape[0], bit_depth, C.c_void_p(out.ctypes.data)),
                    "vtmme_mctf_apply_motion")
        return out

    def mctf_bilateral(self, org_id, corr_ids, weights, width, height, bit_depth=10):
        """EncTemporalFilter::bilateralFilter for one component: uploaded original plane org_id, uploaded motion-compensated
        neighbours corr_ids, weights float64 [len(corr_ids), 1 << bit_depth] (weight by |ref - org|).  Returns int16
        [height, width]."""
        n = len(corr_ids)
        wt = np.ascontiguousarray(weights, dtype=np.float64)
        assert wt.shape == (n, 1 << bit_depth)
        out = np.zeros((height, width), np.int16)
        self._check(self.L.vtmme_mctf_bilateral(self.ctx, org_id, n, (C.c_int32 * n)(*corr_ids), C.c_void_p(wt.ctypes.data),
                                                bit_depth, C.c_void_p(out.ctypes.data)), "vtmme_mctf_bilateral")
        return out

    # ---- decoder-side MV refinement ------------------------------------------------------------------------
    def dmvr_refine(self, ref_pic0, ref_pic1, blocks, bit_depth=10, max_cu=128):
        """The search of InterPrediction::xProcessDMVR (CommonLib/InterPrediction.cpp:2098-2154) for a batch of sub-blocks.
        blocks: int32 [n, 8] {x, y, w, h, mvL0x, mvL0y, mvL1x, mvL1y} (MVs in 1/16 sample).
        Returns int32 [n, 4] {mvdL0SubPu.hor, .ver, minCost, notZeroCost}."""
        blk = np.ascontiguousarray(np.asarray(blocks, dtype=np.int32).reshape(-1, 8))
        out = np.zeros((len(blk), 4), np.int32)
        self._check(self.L.vtmme_dmvr_refine(self.ctx, ref_pic0, ref_pic1, bit_depth, max_cu, len(blk),
                                             C.cast(blk.ctypes.data, C.POINTER(CDmvrBlock)),
                                             C.cast(out.ctypes.data, C.POINTER(CDmvrResult))), "vtmme_dmvr_refine")
        return out

    def dmvr_final_mc(self, comp, ref_pic, blocks, bit_depth=10, max_cu=128):
        """InterPrediction::xFinalPaddedMCForDMVR for a batch of sub-blocks of one list: blocks int32 [n, 8] {x, y, w, h (luma),
        merge MV x, y, refined MV x, y}.  Returns the packed int16 14-bit predictions (comp 1: the 4:2:0 chroma plane)."""
        blk = np.ascontiguousarray(np.asarray(blocks, dtype=np.int32).reshape(-1, 8))
        cs = 1 if comp else 0
        total = int(((blk[:, 2] >> cs) * (blk[:, 3] >> cs)).sum())
        out = np.zeros(total, np.int16)
        self._check(self.L.vtmme_dmvr_final_mc(self.ctx, comp, ref_pic, bit_depth, max_cu, len(blk),
                                               C.cast(blk.ctypes.data, C.POINTER(CDmvrBlock)), C.c_void_p(out.ctypes.data)),
                    "vtmme_dmvr_final_mc")
        return out

    # ---- affine ME primitives ----------------------------------------------------------------------------
    def affine_sobel(self, vertical, pred):
        """AffineGradientSearch's Sobel entries on one prediction block (int16 2-D, views ok) -> int32 [h, w]."""
        assert pred.dtype == np.int16
        h, w = pred.shape
        out = np.zeros((h, w), np.int32)
        self._check(self.L.vtmme_affine_sobel_host(self.ctx, vertical, C.c_void_p(pred.ctypes.data), pred.strides[0] // 2, w, h,
                                                   C.c_void_p(out.ctypes.data), w), "vtmme_affine_sobel_host")
        return out

    def affine_equal_coeff(self, residue, d0, d1, six, coeff=None):
        """xEqualCoeffComputer: residue int16 [h, w], derivatives int32 [h, w]; accumulates into coeff (int64 [7, 7])."""
        h, w = residue.shape
        if coeff is None:
            coeff = np.zeros((7, 7), np.int64)
        self._check(self.L.vtmme_affine_equal_coeff_host(self.ctx, C.c_void_p(residue.ctypes.data), residue.strides[0] // 2,
                                                         C.c_void_p(d0.ctypes.data), C.c_void_p(d1.ctypes.data), d0.strides[0] // 4, w, h,
                                                         int(six), C.c_void_p(coeff.ctypes.data)), "vtmme_affine_equal_coeff_host")
        return coeff

    def affine_gradient_step(self, blocks):
        """blocks: list of (org, pred, sixParam) with int16 2-D arrays of equal shape.  Returns int64 [n, 7, 7]."""
        n = len(blocks)
        arr = (CAffineBlock * n)()
        for i, (o, p, six) in enumerate(blocks):
            assert o.dtype == np.int16 and p.dtype == np.int16 and o.shape == p.shape
            arr[i] = CAffineBlock(o.ctypes.data, o.strides[0] // 2, p.ctypes.data, p.strides[0] // 2, o.shape[1], o.shape[0], int(six), 0)
        out = np.zeros((n, 7, 7), np.int64)
        self._check(self.L.vtmme_affine_gradient_step(self.ctx, n, arr, C.c_void_p(out.ctypes.data)), "vtmme_affine_gradient_step")
        return out

    # ---- symmetric-MVD search -----------------------------------------------------------------------------
    def smvd_search(self, jobs):
        """InterSearch::xSymmetricMotionEstimation for a batch of PUs.  jobs: list of dicts with the fields of vtmme_smvd
        (org: optional int16 2-D array).  Returns a list of (curMvX, curMvY, tarMvX, tarMvY, cost)."""
        n = len(jobs)
        arr = (CSmvd * n)()
        keep = []
        for i, j in enumerate(jobs):
            org = j.get("org")
            if org is not None:
                org = np.ascontiguousarray(org, dtype=np.int16)
                keep.append(org)
            arr[i] = CSmvd(int(j.get("curPic", 0)), int(j["refPicCur"]), int(j["refPicTar"]), int(j["x"]), int(j["y"]), int(j["w"]),
                           int(j["h"]), org.ctypes.data if org is not None else None, org.shape[1] if org is not None else 0,
                           int(j.get("maxCu", 128)), int(j.get("bitDepth", 10)), int(j["imv"]), int(j["curPred"][0]),
                           int(j["curPred"][1]), int(j["tarPred"][0]), int(j["tarPred"][1]), int(j["curMv"][0]), int(j["curMv"][1]),
                           int(j["tarMv"][0]), int(j["tarMv"][1]), int(j.get("clipBiPred", 0)), int(j.get("useHad", 1)),
                           int(j.get("bcwIdx", 2)), float(j["lambdaMotion"]), int(j["cost"]))
        res = (CSmvdResult * n)()
        self._check(self.L.vtmme_smvd_search(self.ctx, n, arr, res), "vtmme_smvd_search")
        return [(r.curMvX, r.curMvY, r.tarMvX, r.tarMvY, int(r.cost)) for r in res]

    # ---- candidate distortion (AMVP template cost / ME seeds) ----------------------------------------------
    def cand_sad(self, jobs, bit_depth=10, use_alt_hpel=0):
        """jobs: list of dicts {curPic, refPic, x, y, w, h, mv: [(mvX, mvY), ...] in 1/16 sample, subShift=0, org=None}
        (org: optional int16 2-D array replacing the block read from curPic).  Returns a list of per-job SAD lists."""
        arr = (CCandJob * len(jobs))()
        keep = []
        total = 0
        for i, j in enumerate(jobs):
            mv = np.ascontiguousarray(np.asarray(j["mv"], dtype=np.int32).reshape(-1, 2))
            org = j.get("org")
            keep.append((mv, org))
            arr[i] = CCandJob(int(j.get("curPic", 0)), int(j["refPic"]), int(j["x"]), int(j["y"]), int(j["w"]), int(j["h"]),
                              C.c_void_p(org.ctypes.data) if org is not None else None,
                              org.strides[0] // 2 if org is not None else 0, len(mv), C.c_void_p(mv.ctypes.data),
                              int(j.get("subShift", 0)), 0)
            total += len(mv)
        out = (C.c_uint64 * total)()
        self._check(self.L.vtmme_cand_sad(self.ctx, bit_depth, use_alt_hpel, len(jobs), arr, out), "vtmme_cand_sad")
        res, k = [], 0
        for mv, _ in keep:
            res.append([int(out[k + c]) for c in range(len(mv))])
            k += len(mv)
        return res
