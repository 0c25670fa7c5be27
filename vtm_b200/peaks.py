"""INT/ALU-pipe peak of the device, measured with the library's own microbenchmark (vtmme_int_peak)."""
import ctypes as C

from .lib import VtmmeError, load_library

VARIANTS = ["vabsdiff", "iadd3", "imad", "lop3", "prmt", "vabsdiff+imad", "fadd_abs(denormal)", "vabsdiff:fadd 2:1",
            "viadd.16x2", "viaddmnmx.s16x2", "vabsdiff4.u8", "vabsdiff:fadd 1:1", "idp.2a (dp2a)", "idp.4a (dp4a)",
            "fadd2 packed |a-b| (px/clk/SM)", "vabsdiff:fadd2 8:8 (px/clk/SM)", "vabsdiff:fadd2 6:10 (px/clk/SM)",
            "vabsdiff:fadd2 10:6 (px/clk/SM)", "vabsdiff:fadd2 12:4 (px/clk/SM)", "vabsdiff:fadd honest 12:4",
            "vabsdiff:fadd honest 10:6", "vabsdiff:fadd honest 8:8"]


def int_peak(variant, iters=4096):
    """-> dict(lane_instr_per_clk_per_sm, ms, sm_mhz) for one instruction class."""
    L = load_library()
    rate, ms, mhz = C.c_double(), C.c_double(), C.c_double()
    rc = L.vtmme_int_peak(variant, iters, C.byref(rate), C.byref(ms), C.byref(mhz))
    if rc != 0:
        raise VtmmeError("vtmme_int_peak(%d) failed: %d" % (variant, rc))
    return {"variant": VARIANTS[variant], "lane_instr_per_clk_per_sm": rate.value, "ms": ms.value, "sm_mhz": mhz.value}


def all_peaks(iters=4096):
    return [int_peak(v, iters) for v in range(len(VARIANTS))]
