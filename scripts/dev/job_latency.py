"""Per-call latency of vtmme_search (the in-loop encoder's flavour: one job per call, host pattern, sync on return)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vtm_b200
from vtm_b200 import Job
from tests.helpers import pad_plane, MARGIN
rng = np.random.default_rng(1)
W, H = 832, 480
ref = rng.integers(0, 1024, (H, W), dtype=np.int16)
cur = np.ascontiguousarray(np.roll(ref, (3, -5), (0, 1)))
ms = vtm_b200.MotionSearch(0)
ms.upload_picture(1, cur)
ms.upload_picture(2, pad_plane(ref), MARGIN)
for (w, h, sr) in [(8, 8, 64), (4, 8, 64), (16, 16, 64), (32, 32, 64), (64, 64, 64), (128, 128, 64), (16, 16, 4), (16, 16, 128)]:
    org = np.ascontiguousarray(cur[200:200 + h, 300:300 + w])
    j = Job(1, 2, 300, 200, w, h, (-sr, sr, -sr, sr), (0, 0), 0, 1 if h > 8 and w <= 64 else 0, 10, 1, 0, 1, 31.33, org)
    for _ in range(20):
        ms.search([j])
    t = time.perf_counter()
    n = 300
    for _ in range(n):
        ms.search([j])
    dt = (time.perf_counter() - t) / n
    j0 = Job(1, 2, 300, 200, w, h, (-sr, sr, -sr, sr), (0, 0), 0, 1 if h > 8 and w <= 64 else 0, 10, 1, 0, 0, 31.33, org)
    for _ in range(5):
        ms.search([j0])
    t = time.perf_counter()
    for _ in range(n):
        ms.search([j0])
    dt0 = (time.perf_counter() - t) / n
    print("%3dx%-3d SR=%3d : %.1f us per call, %.1f us integer-only (python+ctypes overhead included)" % (w, h, sr, dt * 1e6, dt0 * 1e6))
