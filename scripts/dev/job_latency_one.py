"""One shape per process so the library's VTMME_TIMING summary is per shape: job_latency_one.py w h sr frac"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vtm_b200
from vtm_b200 import Job
from tests.helpers import pad_plane, MARGIN
w, h, sr, frac = (int(v) for v in sys.argv[1:5])
rng = np.random.default_rng(1)
W, H = 832, 480
ref = rng.integers(0, 1024, (H, W), dtype=np.int16)
cur = np.ascontiguousarray(np.roll(ref, (3, -5), (0, 1)))
ms = vtm_b200.MotionSearch(0)
ms.upload_picture(1, cur)
ms.upload_picture(2, pad_plane(ref), MARGIN)
org = np.ascontiguousarray(cur[200:200 + h, 300:300 + w])
j = Job(1, 2, 300, 200, w, h, (-sr, sr, -sr, sr), (0, 0), 0, 1 if h > 8 and w <= 64 else 0, 10, 1, 0, frac, 31.33, org)
for _ in range(2000):
    ms.search([j])
print("%dx%d SR=%d frac=%d" % (w, h, sr, frac), file=sys.stderr)
ms.close()
