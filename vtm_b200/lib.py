"""ctypes binding of libvtmme.so (include/vtmme.h).  Fails loudly when the library is missing."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class CSmvd(C.Structure):
    """vtmme_smvd"""
    _fields_ = [("curPic", C.c_int32), ("refPicCur", C.c_int32), ("refPicTar", C.c_int32), ("x", C.c_int32), ("y", C.c_int32),
                ("w", C.c_int32), ("h", C.c_int32), ("org", C.c_void_p), ("orgStride", C.c_int32), ("maxCu", C.c_int32),
                ("bitDepth", C.c_int32), ("imv", C.c_int32), ("curPredX", C.c_int32), ("curPredY", C.c_int32),
                ("tarPredX", C.c_int32), ("tarPredY", C.c_int32), ("curMvX", C.c_int32), ("curMvY", C.c_int32),
                ("tarMvX", C.c_int32), ("tarMvY", C.c_int32), ("clipBiPred", C.c_int32), ("useHad", C.c_int32),
                ("bcwIdx", C.c_int32), ("lambdaMotion", C.c_double), ("cost", C.c_uint64)]


class CAffineBlock(C.Structure):
    """vtmme_affine_block"""
    _fields_ = [("org", C.c_void_p), ("orgStride", C.c_int32), ("pred", C.c_void_p), ("predStride", C.c_int32), ("w", C.c_int32),
                ("h", C.c_int32), ("sixParam", C.c_int32), ("reserved", C.c_int32)]


class CSmvdResult(C.Structure):
    """vtmme_smvd_result"""
    _fields_ = [("curMvX", C.c_int32), ("curMvY", C.c_int32), ("tarMvX", C.c_int32), ("tarMvY", C.c_int32), ("cost", C.c_uint64)]


class VtmmeError(RuntimeError):
    pass


ERR_NAMES = {0: "VTMME_OK", -1: "VTMME_ERR_CUDA", -2: "VTMME_ERR_ARG", -3: "VTMME_ERR_NOMEM",
             -4: "VTMME_ERR_NOPIC", -5: "VTMME_ERR_RANGE"}

# every symbol include/vtmme.h declares (tests check the built library exports all of them)
SYMBOLS = ["vtmme_create", "vtmme_destroy", "vtmme_last_error", "vtmme_set_stream", "vtmme_synchronize",
           "vtmme_launch_count", "vtmme_set_profiling", "vtmme_frame_kernel_ms", "vtmme_upload_picture", "vtmme_upload_picture_async", "vtmme_upload_picture_device", "vtmme_release_picture",
           "vtmme_search", "vtmme_frame_cu_count", "vtmme_search_frames", "vtmme_search_frames_device",
           "vtmme_dist_batch", "vtmme_dist_host", "vtmme_interp_batch", "vtmme_interp_host", "vtmme_filter_host",
           "vtmme_mc_batch", "vtmme_mc_host", "vtmme_add_avg", "vtmme_remove_high_freq", "vtmme_add_weighted_avg", "vtmme_remove_weight_high_freq", "vtmme_cand_sad", "vtmme_dmvr_refine", "vtmme_dmvr_final_mc", "vtmme_smvd_search", "vtmme_affine_sobel_host", "vtmme_affine_equal_coeff_host", "vtmme_affine_gradient_step", "vtmme_mctf_me", "vtmme_mctf_apply_motion", "vtmme_mctf_bilateral", "vtmme_int_peak"]


class CAmvr(C.Structure):
    """vtmme_amvr"""
    _fields_ = [("imv", C.c_int32), ("numCand", C.c_int32), ("candX", C.c_int32 * 2), ("candY", C.c_int32 * 2),
                ("mvpIdx", C.c_int32), ("mvpIdxBits", C.c_uint32 * 2), ("bits", C.c_uint32),
                ("picW", C.c_int32), ("picH", C.c_int32), ("maxCuW", C.c_int32), ("maxCuH", C.c_int32),
                ("fWeight", C.c_double)]


class CDmvrBlock(C.Structure):
    """vtmme_dmvr_block"""
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("mvL0x", C.c_int32),
                ("mvL0y", C.c_int32), ("mvL1x", C.c_int32), ("mvL1y", C.c_int32)]


class CDmvrResult(C.Structure):
    """vtmme_dmvr_result"""
    _fields_ = [("mvdX", C.c_int32), ("mvdY", C.c_int32), ("minCost", C.c_uint32), ("notZeroCost", C.c_int32)]


class CTz(C.Structure):
    """vtmme_tz"""
    _fields_ = [("startX", C.c_int32), ("startY", C.c_int32), ("hasInt2Nx2N", C.c_int32), ("int2Nx2NX", C.c_int32),
                ("int2Nx2NY", C.c_int32), ("nSeeds", C.c_int32), ("seedX", C.c_int32 * 16), ("seedY", C.c_int32 * 16),
                ("searchRange", C.c_int32), ("extended", C.c_int32), ("fast", C.c_int32), ("firstSearchStop", C.c_int32),
                ("picW", C.c_int32), ("picH", C.c_int32), ("maxCu", C.c_int32), ("selective", C.c_int32),
                ("stagedSad", C.c_int32)]


class CJob(C.Structure):
    """vtmme_job"""
    _fields_ = [("curPic", C.c_int32), ("refPic", C.c_int32), ("x", C.c_int32), ("y", C.c_int32),
                ("w", C.c_int32), ("h", C.c_int32), ("org", C.c_void_p), ("orgStride", C.c_int32),
                ("srLeft", C.c_int32), ("srRight", C.c_int32), ("srTop", C.c_int32), ("srBottom", C.c_int32),
                ("predQx", C.c_int32), ("predQy", C.c_int32), ("imvShift", C.c_int32), ("subShift", C.c_int32),
                ("bitDepth", C.c_int32), ("useHad", C.c_int32), ("useAltHpel", C.c_int32), ("fracMode", C.c_int32),
                ("lambdaMotion", C.c_double), ("amvr", C.POINTER(CAmvr)), ("tz", C.POINTER(CTz))]


class CResult(C.Structure):
    """vtmme_result"""
    _fields_ = [("mvX", C.c_int32), ("mvY", C.c_int32), ("intSad", C.c_uint64),
                ("halfX", C.c_int32), ("halfY", C.c_int32), ("qterX", C.c_int32), ("qterY", C.c_int32),
                ("fracCost", C.c_uint64),
                ("amvrMvX", C.c_int32), ("amvrMvY", C.c_int32), ("mvpIdx", C.c_int32), ("bits", C.c_uint32),
                ("cost", C.c_uint64)]

    def tuple(self):
        return (self.mvX, self.mvY, self.intSad, self.halfX, self.halfY, self.qterX, self.qterY, self.fracCost)

    def amvr_tuple(self):
        """xPatternSearchIntRefine's outputs (fracMode 2): (rcMv x, y in 1/16 sample, riMVPIdx, ruiBits, ruiCost)"""
        return (self.amvrMvX, self.amvrMvY, self.mvpIdx, self.bits, self.cost)


class CFrameParams(C.Structure):
    """vtmme_frame_params"""
    _fields_ = [("searchRange", C.c_int32), ("bitDepth", C.c_int32), ("ctuSize", C.c_int32), ("imvShift", C.c_int32),
                ("useHad", C.c_int32), ("fracMode", C.c_int32), ("predSpread", C.c_int32), ("subShiftMode", C.c_int32),
                ("lambdaMotion", C.c_double), ("fastSearch", C.c_int32), ("tzFirstSearchStop", C.c_int32)]


class CMcBlock(C.Structure):
    """vtmme_mc_block"""
    _fields_ = [("refPic", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("mvX", C.c_int32), ("mvY", C.c_int32), ("reserved", C.c_int32)]


class CCandJob(C.Structure):
    """vtmme_cand_job"""
    _fields_ = [("curPic", C.c_int32), ("refPic", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32),
                ("h", C.c_int32), ("org", C.c_void_p), ("orgStride", C.c_int32), ("nCand", C.c_int32),
                ("mv", C.c_void_p), ("subShift", C.c_int32), ("reserved", C.c_int32)]


def library_path():
    return os.path.join(HERE, "libvtmme.so")


def build_library(jobs=8):
    """Compile every CUDA source for sm_100a into vtm_b200/libvtmme.so (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "csrc"), "-j%d" % jobs])
    return library_path()


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise VtmmeError("libvtmme.so is not built (%s): run `make -C vtm_b200/csrc` or __graft_entry__.build(); "
                         "there is no CPU fallback for the motion-search path" % path)
    L = C.CDLL(path)
    P, I, I64 = C.c_void_p, C.c_int, C.c_int64
    L.vtmme_create.argtypes = [I, C.POINTER(P)]
    L.vtmme_destroy.argtypes = [P]
    L.vtmme_destroy.restype = None
    L.vtmme_last_error.argtypes = [P]
    L.vtmme_last_error.restype = C.c_char_p
    L.vtmme_set_stream.argtypes = [P, P]
    L.vtmme_synchronize.argtypes = [P]
    L.vtmme_launch_count.argtypes = [P]
    L.vtmme_launch_count.restype = C.c_uint64
    L.vtmme_set_profiling.argtypes = [P, I]
    L.vtmme_frame_kernel_ms.argtypes = [P, C.POINTER(C.c_float)]
    L.vtmme_upload_picture.argtypes = [P, I, P, I, I, I, I, I]
    L.vtmme_upload_picture_async.argtypes = [P, I, P, I, I, I, I, I]
    L.vtmme_upload_picture_device.argtypes = [P, I, P, I, I, I, I, I]
    L.vtmme_release_picture.argtypes = [P, I]
    L.vtmme_search.argtypes = [P, C.POINTER(CJob), I, C.POINTER(CResult)]
    L.vtmme_frame_cu_count.argtypes = [I, I, C.POINTER(C.c_int32)]
    L.vtmme_search_frames.argtypes = [P, I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(CFrameParams), P, P]
    L.vtmme_search_frames_device.argtypes = [P, I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(CFrameParams), P, P]
    L.vtmme_dist_batch.argtypes = [P, I, P, I, I64, P, I, I64, I, I, I, I, P]
    L.vtmme_dist_host.argtypes = [P, I, P, I, P, I, I, I, I, C.POINTER(C.c_uint64)]
    L.vtmme_interp_batch.argtypes = [P, I, I, P, I, I64, P, I, I64, I, I, I, I, I, I, I, I]
    L.vtmme_interp_host.argtypes = [P, I, I, P, I, P, I, I, I, I, I, I, I, I]
    L.vtmme_filter_host.argtypes = [P, I, I, I, I, I, P, I, P, I, I, I, P, I]
    L.vtmme_mc_batch.argtypes = [P, I, I, I, I, I, C.POINTER(CMcBlock), P]
    L.vtmme_mc_host.argtypes = [P, I, I, I, I, I, C.POINTER(CMcBlock), P]
    L.vtmme_add_avg.argtypes = [P, P, P, P, I64, I]
    L.vtmme_remove_high_freq.argtypes = [P, P, P, I64, I, I]
    L.vtmme_add_weighted_avg.argtypes = [P, P, P, P, I64, I, I]
    L.vtmme_remove_weight_high_freq.argtypes = [P, P, P, I64, I, I, I]
    L.vtmme_cand_sad.argtypes = [P, I, I, I, C.POINTER(CCandJob), C.POINTER(C.c_uint64)]
    L.vtmme_dmvr_refine.argtypes = [P, I, I, I, I, I, C.POINTER(CDmvrBlock), C.POINTER(CDmvrResult)]
    L.vtmme_dmvr_final_mc.argtypes = [P, I, I, I, I, I, C.POINTER(CDmvrBlock), P]
    L.vtmme_smvd_search.argtypes = [P, I, C.POINTER(CSmvd), C.POINTER(CSmvdResult)]
    L.vtmme_affine_sobel_host.argtypes = [P, I, P, I, I, I, P, I]
    L.vtmme_affine_equal_coeff_host.argtypes = [P, P, I, P, P, I, I, I, I, P]
    L.vtmme_affine_gradient_step.argtypes = [P, I, C.POINTER(CAffineBlock), P]
    L.vtmme_mctf_me.argtypes = [P, I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), I, P]
    L.vtmme_mctf_apply_motion.argtypes = [P, I, I, I, P, I, I, I, P]
    L.vtmme_mctf_bilateral.argtypes = [P, I, I, C.POINTER(C.c_int32), P, I, P]
    L.vtmme_int_peak.argtypes = [I, I, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _lib = L
    return L
