"""Generates tests/golden/mctf_golden.npz from the UNMODIFIED reference compiled here (oracle/_ref/libvtmref.so):
EncTemporalFilter::motionEstimation on one synthetic 160x96 picture pair.  Run in the build container only:

    python tests/golden/make_golden_mctf.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from tests.helpers import pad_plane  # noqa: E402


def main():
    from vtm_b200.synth import make_pair
    R = B.ref()
    assert R is not None, "build the reference first: make -f oracle/Makefile.ref -j8 all"
    w, h = 160, 96
    cur, ref, _ = make_pair(888, w, h, max_global=8, max_local=12, n_rects=2, sigma=5.0)
    curp, refp = pad_plane(cur, 128), pad_plane(ref, 128)
    stride = curp.shape[1]
    off = 128 * stride + 128
    mv = np.zeros((h // 4, w // 4, 3), np.int32)
    R.ref_mctf_me(B.ptr(curp, off), stride, B.ptr(refp, off), stride, w, h, 10, C.c_void_p(mv.ctypes.data))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mctf_golden.npz")
    np.savez_compressed(path, cur=np.ascontiguousarray(cur), ref=np.ascontiguousarray(ref), mv=mv)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
