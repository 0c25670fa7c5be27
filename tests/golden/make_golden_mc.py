"""Generates tests/golden/mc_golden.npz: motion-compensated predictions produced by the UNMODIFIED reference's
InterPrediction::xPredInterBlk (through oracle/_ref/libvtmref.so, see oracle/ref_harness.cpp ref_mc_blocks), plus
addAvg / removeHighFreq outputs.  Run in the build container only:

    python tests/golden/make_golden_mc.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from tests.helpers import mc_cases  # noqa: E402

W, H, M = 128, 64, 48   # luma size; chroma planes are W/2 x H/2


def main():
    R = B.ref()
    assert R is not None, "build the reference first: make -f oracle/Makefile.ref -j8 all"
    rng = np.random.default_rng(20261019)
    out = {"dims": np.array([W, H, M])}
    for comp, name in ((0, "y"), (1, "c")):
        cw, ch = (W, H) if comp == 0 else (W // 2, H // 2)
        plane = rng.integers(0, 1024, (ch, cw), dtype=np.int16)
        padded = np.ascontiguousarray(np.pad(plane, M, mode="edge"))
        out["plane_" + name] = plane
        blks = mc_cases(77 + comp, comp, cw, ch, 90, sizes=[4, 8, 16, 32, 64] if comp == 0 else [2, 4, 8, 16, 32], max_mv_pel=20)
        ba = np.array(blks, dtype=np.int32)
        out["blocks_" + name] = ba
        n = int((ba[:, 2] * ba[:, 3]).sum())
        for bi in (0, 1):
            for alt in (0, 1):
                if comp == 1 and alt:
                    continue
                dst = np.zeros(n, np.int16)
                rc = R.ref_mc_blocks(comp, B.ptr(padded), padded.shape[1], W, H, M, len(blks), C.c_void_p(ba.ctypes.data), bi, 10,
                                     alt, B.ptr(dst), None)
                assert rc == 0
                out["pred_%s_bi%d_alt%d" % (name, bi, alt)] = dst
    # bi-prediction helpers on 64x32 samples
    s0 = rng.integers(-8192, 8192, (32, 64), dtype=np.int16)
    s1 = rng.integers(-8192, 8192, (32, 64), dtype=np.int16)
    avg = np.zeros((32, 64), np.int16)
    R.ref_add_avg(B.ptr(s0), B.ptr(s1), B.ptr(avg), 64, 32, 10)
    org = rng.integers(0, 1024, (32, 64), dtype=np.int16)
    pred = rng.integers(0, 1024, (32, 64), dtype=np.int16)
    hf = []
    for clip in (0, 1):
        t = org.copy()
        R.ref_remove_high_freq(B.ptr(t), 64, B.ptr(pred), 64, 64, 32, clip, 10)
        hf.append(t)
    out.update(avg_s0=s0, avg_s1=s1, avg_out=avg, hf_org=org, hf_pred=pred, hf_out0=hf[0], hf_out1=hf[1])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mc_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
