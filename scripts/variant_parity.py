"""Development aid: parity of one tree-kernel variant (VTMME_TREE_VARIANT) against the oracle, per CU level."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vtm_b200  # noqa: E402
from oracle import bindings as B  # noqa: E402
from tests.helpers import MARGIN, gpu_tuple, oracle_frame_search, pad_plane  # noqa: E402
from vtm_b200 import FrameParams  # noqa: E402
from vtm_b200.synth import make_pair  # noqa: E402

L = B.oracle()
ms = vtm_b200.MotionSearch(0)
w, h, sr, lam = 256, 128, 16, 31.33
cur, ref, _ = make_pair(1, w, h, max_global=12, max_local=14, n_rects=3, sigma=4.0)
refp = pad_plane(ref)
ms.upload_picture(1, cur)
ms.upload_picture(2, refp, MARGIN)
ncu = ms.set_frame_size(w, h)
got = ms.search_frames([1], [2], FrameParams(searchRange=sr, lambdaMotion=lam, fracMode=0))
want = oracle_frame_search(L, cur, refp, MARGIN, sr, lam, None, 0)
off = ms._off
for l in range(5):
    idx = range(off[l], off[l + 1])
    bad = [i for i in idx if gpu_tuple(got[0][i]) != want[i]]
    print("variant", os.environ.get("VTMME_TREE_VARIANT"), "level", l, "bad", len(bad), "of", len(idx),
          [(gpu_tuple(got[0][i]), want[i]) for i in bad[:2]])
