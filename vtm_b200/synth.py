"""Synthetic luma frame pairs for the standalone batched-ME benchmark (BASELINE config 4).

Pair p (seed 1000+p): box-blurred uniform noise texture on a canvas larger than the picture; `ref` is a crop of
the canvas, `cur` is the crop displaced by a global motion with a quarter-pel phase (rendered with the VVC
8-tap luma filter), eight rectangles carry their own integer motion, Gaussian noise (sigma 8) on top,
clipped to 10 bit.  JVET test sequences are not available offline.

numpy version: tests (small sizes, CPU).  torch version: the benchmark (generated on the GPU).
"""
import numpy as np

LUMA_FILTER_QPEL = {
    0: (0, 0, 0, 64, 0, 0, 0, 0),
    1: (-1, 4, -10, 58, 17, -5, 1, 0),
    2: (-1, 4, -11, 40, 40, -11, 4, -1),
    3: (0, 1, -5, 17, 58, -10, 4, -1),
}
CANVAS_PAD = 128


def _subpel_np(img, fx, fy):
    """img int32 [H,W]; returns the block sampled at +fx/4, +fy/4 (valid region shrinks by the 8-tap support)."""
    a = img.astype(np.int64)
    if fx:
        c = LUMA_FILTER_QPEL[fx]
        acc = np.zeros_like(a[:, 3:-4])
        for k in range(8):
            acc += c[k] * a[:, k:a.shape[1] - 7 + k]
        a = np.pad((acc + 32) >> 6, ((0, 0), (3, 4)), mode="edge")
    if fy:
        c = LUMA_FILTER_QPEL[fy]
        acc = np.zeros_like(a[3:-4, :])
        for k in range(8):
            acc += c[k] * a[k:a.shape[0] - 7 + k, :]
        a = np.pad((acc + 32) >> 6, ((3, 4), (0, 0)), mode="edge")
    return a


def make_pair(p, width=1920, height=1080, max_global=48, max_local=60, n_rects=8, sigma=8.0, bit_depth=10):
    """-> (cur, ref, info): int16 [height,width] planes (no border) and the planted motion."""
    rng = np.random.default_rng(1000 + p)
    maxv = (1 << bit_depth) - 1
    pad = CANVAS_PAD
    H, W = height + 2 * pad, width + 2 * pad
    noise = rng.integers(0, 1 << bit_depth, (H + 2, W + 2)).astype(np.int64)
    canvas = sum(noise[dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3)) // 9
    ref = canvas[pad:pad + height, pad:pad + width]
    gx, gy = (int(v) for v in rng.integers(-max_global, max_global + 1, 2))
    fx, fy = (int(v) for v in rng.integers(0, 4, 2))
    shifted = _subpel_np(canvas, fx, fy)
    cur = shifted[pad + gy:pad + gy + height, pad + gx:pad + gx + width].copy()
    rects = []
    for _ in range(n_rects):
        rw, rh = (int(v) for v in rng.integers(32, 257, 2))
        rw, rh = min(rw, width), min(rh, height)
        rx, ry = int(rng.integers(0, width - rw + 1)), int(rng.integers(0, height - rh + 1))
        mx, my = (int(v) for v in rng.integers(-max_local, max_local + 1, 2))
        cur[ry:ry + rh, rx:rx + rw] = canvas[pad + ry + my:pad + ry + my + rh, pad + rx + mx:pad + rx + mx + rw]
        rects.append((rx, ry, rw, rh, mx, my))
    cur = cur + np.rint(rng.normal(0.0, sigma, cur.shape)).astype(np.int64)
    cur = np.clip(cur, 0, maxv).astype(np.int16)
    info = {"global_qpel": (4 * gx + fx, 4 * gy + fy), "rects": rects}
    return np.ascontiguousarray(cur), np.ascontiguousarray(ref.astype(np.int16)), info


def random_predictors(p, n_cu, spread_px=16):
    """Seeded random quarter-pel predictors within +-spread_px (run B of config 4): int16 [n_cu, 2]."""
    rng = np.random.default_rng(5000 + p)
    return rng.integers(-4 * spread_px, 4 * spread_px + 1, (n_cu, 2)).astype(np.int16)


def make_pairs_torch(ids, device, width=1920, height=1080, max_global=48, max_local=60, n_rects=8, sigma=8.0,
                     bit_depth=10):
    """Same construction on the GPU (torch), one pair per id.  -> cur, ref int16 tensors [n,height,width]."""
    import torch
    maxv = (1 << bit_depth) - 1
    pad = CANVAS_PAD
    H, W = height + 2 * pad, width + 2 * pad
    curs, refs = [], []
    for p in ids:
        g = torch.Generator(device=device)
        g.manual_seed(1000 + int(p))
        host = np.random.default_rng(1000 + int(p))
        noise = torch.randint(0, 1 << bit_depth, (H + 2, W + 2), generator=g, device=device, dtype=torch.int32)
        canvas = sum(noise[dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3)) // 9
        gx, gy = (int(v) for v in host.integers(-max_global, max_global + 1, 2))
        fx, fy = (int(v) for v in host.integers(0, 4, 2))
        a = canvas
        if fx:
            c = LUMA_FILTER_QPEL[fx]
            acc = sum(c[k] * a[:, k:a.shape[1] - 7 + k] for k in range(8))
            a = torch.nn.functional.pad(((acc + 32) >> 6)[None, None].float(), (3, 4, 0, 0), mode="replicate")[0, 0].to(torch.int32)
        if fy:
            c = LUMA_FILTER_QPEL[fy]
            acc = sum(c[k] * a[k:a.shape[0] - 7 + k, :] for k in range(8))
            a = torch.nn.functional.pad(((acc + 32) >> 6)[None, None].float(), (0, 0, 3, 4), mode="replicate")[0, 0].to(torch.int32)
        cur = a[pad + gy:pad + gy + height, pad + gx:pad + gx + width].clone()
        for _ in range(n_rects):
            rw, rh = (int(v) for v in host.integers(32, 257, 2))
            rw, rh = min(rw, width), min(rh, height)
            rx, ry = int(host.integers(0, width - rw + 1)), int(host.integers(0, height - rh + 1))
            mx, my = (int(v) for v in host.integers(-max_local, max_local + 1, 2))
            cur[ry:ry + rh, rx:rx + rw] = canvas[pad + ry + my:pad + ry + my + rh, pad + rx + mx:pad + rx + mx + rw]
        nz = torch.randn(cur.shape, generator=g, device=device) * sigma
        cur = (cur + torch.round(nz).to(torch.int32)).clamp_(0, maxv).to(torch.int16)
        curs.append(cur)
        refs.append(canvas[pad:pad + height, pad:pad + width].to(torch.int16))
    return torch.stack(curs), torch.stack(refs)
