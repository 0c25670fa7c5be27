"""Per-kernel time of the frame search for one mode: frame_time.py <A|B|FEN> [pairs]  (VTMME_TREE_VARIANT selects the kernel)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import vtm_b200
from vtm_b200 import FrameParams
from vtm_b200.synth import make_pairs_torch, random_predictors
mode = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ms = vtm_b200.MotionSearch(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ms.set_stream(st.cuda_stream)
W, H = 1920, 1080
cur, ref = make_pairs_torch(range(n), "cuda", W, H)
for i in range(n):
    ms.upload_picture_device(2 * i, cur[i].data_ptr(), W, W, H)
    ms.upload_picture_device(2 * i + 1, ref[i].data_ptr(), W, W, H)
ncu = ms.set_frame_size(W, H)
spread = 16 if mode == "B" else 0
prm = FrameParams(searchRange=64, lambdaMotion=31.33, predSpread=2 * spread + 1 if spread else 0, subShiftMode=2 if mode == "FEN" else 0)
pred = None
if spread:
    pred = np.stack([random_predictors(77 + i, ncu, spread) for i in range(n)])
ms.set_profiling(True)
best = None
for it in range(4):
    ms.search_frames(list(range(0, 2 * n, 2)), list(range(1, 2 * n, 2)), prm, pred)
    k = ms.frame_kernel_ms()
    best = k if best is None or k[0] < best[0] else best
print("mode %s variant %s: tree %.3f upper %.3f frac %.3f ms/pair" % (mode, os.environ.get("VTMME_TREE_VARIANT", "default"), best[0] / n, best[1] / n, best[2] / n))
