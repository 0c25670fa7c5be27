"""TEST INFRASTRUCTURE — ctypes bindings for the CPU oracle and the compiled reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
The product package (vtm_b200) never does.

  oracle()  -> libvtmoracle.so   plain-C restatement (oracle/vtm_oracle.c), built on demand with gcc
  ref()     -> oracle/_ref/libvtmref.so  the UNMODIFIED reference behind a C shim (oracle/ref_harness.cpp),
               built by oracle/Makefile.ref where /root/reference exists; None when it was never built.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class Job(C.Structure):
    """vo_job / RefSearchJob (identical layout)."""
    _fields_ = [("org", C.c_void_p), ("orgStride", C.c_int), ("w", C.c_int), ("h", C.c_int),
                ("refAtPU", C.c_void_p), ("refStride", C.c_int),
                ("srLeft", C.c_int), ("srRight", C.c_int), ("srTop", C.c_int), ("srBottom", C.c_int),
                ("predQx", C.c_int), ("predQy", C.c_int), ("imvShift", C.c_int), ("subShiftMode", C.c_int),
                ("bitDepth", C.c_int), ("useHad", C.c_int), ("useAltHpel", C.c_int), ("doFrac", C.c_int),
                ("lambdaMotion", C.c_double)]


class Result(C.Structure):
    """vo_result / RefSearchResult (identical layout)."""
    _fields_ = [("mvX", C.c_int), ("mvY", C.c_int), ("intSad", C.c_uint64),
                ("halfX", C.c_int), ("halfY", C.c_int), ("qterX", C.c_int), ("qterY", C.c_int),
                ("fracCost", C.c_uint64)]

    def tuple(self):
        return (self.mvX, self.mvY, self.intSad, self.halfX, self.halfY, self.qterX, self.qterY, self.fracCost)


class IntRefine(C.Structure):
    """vo_int_refine_io / RefIntRefine (identical layout): xPatternSearchIntRefine's in/out state."""
    _fields_ = [("imv", C.c_int), ("mvX", C.c_int), ("mvY", C.c_int), ("numCand", C.c_int),
                ("candX", C.c_int * 2), ("candY", C.c_int * 2), ("mvpIdx", C.c_int),
                ("mvpIdxBits", C.c_uint32 * 2), ("bits", C.c_uint32), ("fWeight", C.c_double),
                ("posX", C.c_int), ("posY", C.c_int), ("picW", C.c_int), ("picH", C.c_int),
                ("maxCuW", C.c_int), ("maxCuH", C.c_int), ("cost", C.c_uint64)]

    def tuple(self):
        return (self.mvX, self.mvY, self.mvpIdx, self.bits, self.cost)


class SmvdIo(C.Structure):
    """vo_smvd_io / RefSmvdIo (identical layout): state of an xSymmetricMotionEstimation call."""
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int), ("picW", C.c_int), ("picH", C.c_int),
                ("maxCuW", C.c_int), ("maxCuH", C.c_int), ("bd", C.c_int), ("imv", C.c_int),
                ("curPredX", C.c_int), ("curPredY", C.c_int), ("tarPredX", C.c_int), ("tarPredY", C.c_int),
                ("curMvX", C.c_int), ("curMvY", C.c_int), ("tarMvX", C.c_int), ("tarMvY", C.c_int),
                ("clipBiPred", C.c_int), ("useHad", C.c_int), ("bcwIdx", C.c_int), ("lambda_", C.c_double), ("cost", C.c_uint64)]

    def tuple(self):
        return (self.curMvX, self.curMvY, self.tarMvX, self.tarMvY, self.cost)


class TzParams(C.Structure):
    """vo_tz_params / RefTzParams (identical layout): what xTZSearch receives besides the job."""
    _fields_ = [("startX", C.c_int), ("startY", C.c_int), ("hasInt2Nx2N", C.c_int), ("int2Nx2NX", C.c_int),
                ("int2Nx2NY", C.c_int), ("nSeeds", C.c_int), ("seedX", C.c_int * 16), ("seedY", C.c_int * 16),
                ("searchRange", C.c_int), ("extended", C.c_int), ("fast", C.c_int), ("firstSearchStop", C.c_int),
                ("posX", C.c_int), ("posY", C.c_int), ("picW", C.c_int), ("picH", C.c_int),
                ("maxCuW", C.c_int), ("maxCuH", C.c_int), ("selective", C.c_int)]


def build_oracle():
    subprocess.check_call(["make", "-s", "-f", "oracle/Makefile"], cwd=ROOT)


def build_ref(jobs=8):
    """Compile the reference from /root/reference into oracle/_ref (a few minutes)."""
    subprocess.check_call(["make", "-s", "-f", "oracle/Makefile.ref", "-j%d" % jobs, "all"], cwd=ROOT)


_oracle = None
_ref = None
_P = C.c_void_p
_I = C.c_int


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(HERE, "libvtmoracle.so")
        src = os.path.join(HERE, "vtm_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build_oracle()
        L = C.CDLL(path)
        L.vo_subshift.restype = _I
        L.vo_sad.restype = C.c_uint64
        L.vo_sad.argtypes = [_P, _I, _P, _I, _I, _I, _I]
        L.vo_satd.restype = C.c_uint64
        L.vo_satd.argtypes = [_P, _I, _P, _I, _I, _I]
        L.vo_eg_bits.restype = C.c_uint32
        L.vo_mv_bits.restype = C.c_uint32
        L.vo_mv_cost.restype = C.c_uint64
        L.vo_mv_cost.argtypes = [C.c_double, C.c_uint32]
        L.vo_filter_hor.argtypes = [_I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I]
        L.vo_filter_ver.argtypes = [_I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I]
        L.vo_set_search_range.argtypes = [_I] * 9 + [C.POINTER(_I)] * 4
        L.vo_search.argtypes = [C.POINTER(Job), C.POINTER(Result), _I]
        L.vo_search_batch.argtypes = [C.POINTER(Job), C.POINTER(Result), _I, _I]
        L.vo_search_batch.restype = C.c_double
        L.vo_frac_direct.argtypes = [C.POINTER(Job), _I, _I] + [C.POINTER(_I)] * 4 + [C.POINTER(C.c_uint64)]
        L.vo_pred_qpel.argtypes = [C.POINTER(Job), _I, _I, _I, _I, _I, _P, _I]
        L.vo_me_finish.argtypes = [C.POINTER(Job), C.POINTER(Result), C.c_double, C.c_uint32,
                                   C.POINTER(_I), C.POINTER(_I), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.vo_int_refine.argtypes = [C.POINTER(Job), C.POINTER(IntRefine)]
        L.vo_tz_search.argtypes = [C.POINTER(Job), C.POINTER(TzParams), C.POINTER(_I), C.POINTER(_I),
                                   C.POINTER(C.c_uint64), C.POINTER(_I)]
        L.vo_affine_sobel.restype = None
        L.vo_affine_sobel.argtypes = [_I, _P, _I, _P, _I, _I, _I]
        L.vo_affine_equal_coeff.restype = None
        L.vo_affine_equal_coeff.argtypes = [_P, _I, _P, _P, _I, _P, _I, _I, _I]
        L.vo_mctf_apply_motion.argtypes = [_P, _I, _I, _I, _I, _I, _P, _I, _I, _P, _I]
        L.vo_mctf_me.argtypes = [_P, _I, _P, _I, _I, _I, _I, _P]
        L.vo_mctf_bilateral_weights.restype = None
        L.vo_mctf_bilateral_weights.argtypes = [_I, _I, C.c_double, _I, _I, _I, _P]
        L.vo_mctf_bilateral.restype = None
        L.vo_mctf_bilateral.argtypes = [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, C.c_double, _I, _P, _I]
        L.vo_add_weighted_avg.restype = None
        L.vo_add_weighted_avg.argtypes = [_P, _P, _P, _I, _I, _I]
        L.vo_remove_weight_high_freq.restype = None
        L.vo_remove_weight_high_freq.argtypes = [_P, _P, _I, _I, _I, _I]
        L.vo_smvd_search.restype = None
        L.vo_smvd_search.argtypes = [_P, _I, _P, _P, _I, C.POINTER(SmvdIo)]
        L.vo_dmvr_final_chroma.restype = None
        L.vo_dmvr_final_chroma.argtypes = [_P, _I] + [_I] * 13 + [_P]
        L.vo_dmvr_final_luma.restype = None
        L.vo_dmvr_final_luma.argtypes = [_P, _I] + [_I] * 13 + [_P]
        L.vo_dmvr_block.restype = None
        L.vo_dmvr_block.argtypes = [_P, _P, _I] + [_I] * 13 + [_P]
        L.vo_mctf_error.argtypes = [_P, _I, _P, _I] + [_I] * 7
        L.vo_mctf_error.restype = _I
        L.vo_mc_block.argtypes = [_I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I]
        L.vo_add_avg.argtypes = [_P, _P, _P, _I, _I]
        L.vo_remove_high_freq.argtypes = [_P, _P, _I, _I, _I]
        _oracle = L
    return _oracle


def ref():
    """The compiled reference, or None when oracle/_ref/libvtmref.so does not exist."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libvtmref.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_simd_level.restype = _I
        L.ref_dist.restype = C.c_uint64
        L.ref_dist.argtypes = [_P, _I, _P, _I, _I, _I, _I, _I, _I]
        L.ref_subshift.restype = _I
        L.ref_mv_bits.restype = C.c_uint32
        L.ref_mv_cost.restype = C.c_uint64
        L.ref_mv_cost.argtypes = [C.c_double, C.c_uint32]
        L.ref_filter_hor.argtypes = [_I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I]
        L.ref_filter_ver.argtypes = [_I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I]
        L.ref_search.argtypes = [C.POINTER(Job), C.POINTER(Result)]
        L.ref_affine_sobel.restype = None
        L.ref_affine_sobel.argtypes = [_I, _P, _I, _P, _I, _I, _I]
        L.ref_affine_equal_coeff.restype = None
        L.ref_affine_equal_coeff.argtypes = [_P, _I, _P, _P, _I, _P, _I, _I, _I]
        L.ref_search_batch.argtypes = [C.POINTER(Job), C.POINTER(Result), _I, _I]
        L.ref_search_batch.restype = C.c_double
        L.ref_dist_batch.restype = C.c_double
        L.ref_dist_batch.argtypes = [_P, _P, _I, _I, _I, _I, _I, _I, _P]
        L.ref_filter_batch.restype = C.c_double
        L.ref_filter_batch.argtypes = [_I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I]
        L.ref_int_refine.argtypes = [C.POINTER(Job), C.POINTER(IntRefine)]
        L.ref_tz_search.argtypes = [C.POINTER(Job), C.POINTER(TzParams), C.POINTER(_I), C.POINTER(_I),
                                    C.POINTER(C.c_uint64)]
        L.ref_tz_batch.restype = C.c_double
        L.ref_tz_batch.argtypes = [C.POINTER(Job), C.POINTER(TzParams), _I, _I, _P, _P]
        L.ref_mctf_apply_motion.argtypes = [_P, _P, _I, _I, _I, _P, _P, _P]
        L.ref_mctf_bilateral.restype = None
        L.ref_mctf_bilateral.argtypes = [_P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, C.c_double, _P, _P]
        L.ref_mctf_me.restype = C.c_double
        L.ref_mctf_me.argtypes = [_P, _I, _P, _I, _I, _I, _I, _P]
        L.ref_mc_blocks.argtypes = [_I, _P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P, C.POINTER(C.c_double)]
        L.ref_smvd_search.argtypes = [_P, _I, _P, _P, _I, _I, _I, C.POINTER(SmvdIo)]
        L.ref_dmvr_final.argtypes = [_P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P]
        L.ref_dmvr_final_luma.argtypes = [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P]
        L.ref_dmvr_blocks.argtypes = [_P, _P, _I, _I, _I, _I, _I, _P, _I, _P]
        L.ref_add_avg.argtypes = [_P, _P, _P, _I, _I, _I]
        L.ref_add_weighted_avg.argtypes = [_P, _P, _P, _I, _I, _I, _I]
        L.ref_remove_weight_high_freq.argtypes = [_P, _P, _I, _I, _I, _I, _I]
        L.ref_remove_high_freq.argtypes = [_P, _I, _P, _I, _I, _I, _I, _I]
        _ref = L
    return _ref


def ptr(a, off_elems=0):
    """Address of element `off_elems` of a contiguous int16 array."""
    assert a.dtype == np.int16
    return C.c_void_p(a.ctypes.data + 2 * int(off_elems))


def make_job(org, ref_plane, ref_stride, pu_off, w, h, sr, pred_q, imv_shift=0, sub_shift_mode=0, bit_depth=10,
             use_had=1, use_alt_hpel=0, do_frac=1, lambda_motion=31.33, org_off=0, org_stride=None):
    """org: int16 array holding the original block at element offset org_off (stride org_stride);
    ref_plane: int16 array (padded reference plane), pu_off: element offset of the PU position."""
    if org_stride is None:
        org_stride = w
    return Job(ptr(org, org_off), org_stride, w, h, ptr(ref_plane, pu_off), ref_stride,
               sr[0], sr[1], sr[2], sr[3], pred_q[0], pred_q[1], imv_shift, sub_shift_mode, bit_depth,
               use_had, use_alt_hpel, do_frac, lambda_motion)
