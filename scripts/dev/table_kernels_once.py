"""One launch each of the table-level SATD and 8-tap filter kernels on an HBM-resident batch of 128x128 blocks (after a warm-up
launch) — the program scripts/gpu_r02s.sh profiles with ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vtm_b200  # noqa: E402

ms = vtm_b200.MotionSearch(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ms.set_stream(stream.cuda_stream)
w = h = 128
n = 8192
org = torch.randint(0, 1024, (n, h, w), device="cuda", dtype=torch.int16)
cur = torch.randint(0, 1024, (n, h, w), device="cuda", dtype=torch.int16)
res = torch.zeros(n, dtype=torch.int64, device="cuda")
src = torch.randint(0, 1024, (n, h + 8, w + 8), device="cuda", dtype=torch.int16)
dst = torch.zeros((n, h, w), dtype=torch.int16, device="cuda")
off, ss = 4 * (w + 8) + 4, w + 8
for rep in range(2):
    ms.dist_batch(1, org.data_ptr(), w, w * h, cur.data_ptr(), w, w * h, w, h, 0, n, res.data_ptr())
    ms.interp_batch(0, 0, src.data_ptr() + 2 * off, ss, ss * (h + 8), dst.data_ptr(), w, w * h, w, h, 5, 1, 0, 10, 0, n)
    ms.interp_batch(0, 1, src.data_ptr() + 2 * off, ss, ss * (h + 8), dst.data_ptr(), w, w * h, w, h, 5, 1, 1, 10, 0, n)
    torch.cuda.synchronize()
print("ok", int(res[0]), int(dst[0, 0, 0]))
ms.close()
