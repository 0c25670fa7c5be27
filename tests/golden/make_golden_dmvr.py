"""Generates tests/golden/dmvr_golden.npz from the UNMODIFIED reference compiled here (oracle/_ref/libvtmref.so): the DMVR
search of a sub-block through the reference's own InterPrediction members (xPrefetch, xinitMC, xDMVRCost, xBIPMVRefine,
xDMVRSubPixelErrorSurface; oracle/ref_harness.cpp: ref_dmvr_blocks).  The two reference pictures are the `ref` and `cur`
planes of amvr_tz_golden.npz (true motion between them); a second set reads one picture through both lists (exact matches).
Run in the build container only:

    python tests/golden/make_golden_dmvr.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from tests.helpers import MARGIN, dmvr_cases, pad_plane  # noqa: E402

W, H = 192, 128


def main():
    R = B.ref()
    assert R is not None, "build the reference first: make -f oracle/Makefile.ref -j8 all"
    here = os.path.dirname(os.path.abspath(__file__))
    ga = np.load(os.path.join(here, "amvr_tz_golden.npz"))
    p0, p1 = pad_plane(np.ascontiguousarray(ga["ref"])), pad_plane(np.ascontiguousarray(ga["cur"]))
    stride = p0.shape[1]
    rng = np.random.default_rng(20261021)
    out = {}
    for name, a, b, same in (("pair", p0, p1, False), ("same", p0, p0, True)):
        blk = dmvr_cases(rng, W, H, 160, max_pel=12, same=same)
        res = np.zeros((len(blk), 4), np.int32)
        assert R.ref_dmvr_blocks(B.ptr(a), B.ptr(b), stride, W, H, MARGIN, len(blk), C.c_void_p(blk.ctypes.data), 10,
                                 C.c_void_p(res.ctypes.data)) == 0
        out[name + "_blk"], out[name + "_res"] = blk, res
    path = os.path.join(here, "dmvr_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", sum(len(out[k]) for k in out if k.endswith("_blk")), "cases;",
          "moved", int((out["pair_res"][:, :2] != 0).any(axis=1).sum()), "early exits", int((out["same_res"][:, 3] == 0).sum()))


if __name__ == "__main__":
    main()
