#!/usr/bin/env python
"""Synthetic 4:2:0 test clip for the encoder-level parity runs (BASELINE configs 1-3): textured background with a
global pan, two moving rectangles, Gaussian noise.  JVET sequences are not available offline.

  python integration/make_yuv.py out.yuv --width 416 --height 240 --frames 8 --bits 8
"""
import argparse

import numpy as np


def paste(y, rect, x, y0):
    """rect into y at (x, y0), clipped to the picture (any picture size, rectangles may leave it on every side)"""
    H, W = y.shape
    h, w = rect.shape
    xa, xb, ya, yb = max(x, 0), min(x + w, W), max(y0, 0), min(y0 + h, H)
    if xa < xb and ya < yb:
        y[ya:yb, xa:xb] = rect[ya - y0:yb - y0, xa - x:xb - x]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--width", type=int, default=416)
    ap.add_argument("--height", type=int, default=240)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--bits", type=int, default=8)
    ap.add_argument("--seed", type=int, default=7)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    W, H, maxv = a.width, a.height, (1 << a.bits) - 1
    pad = 96
    base = rng.integers(0, maxv + 1, (H + 2 * pad + 4, W + 2 * pad + 4)).astype(np.float64)
    k = np.ones(5) / 5.0
    for ax in (0, 1):
        base = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), ax, base)
    base = (base - base.min()) / (base.max() - base.min()) * maxv
    yy, xx = np.mgrid[0:base.shape[0], 0:base.shape[1]]
    base = 0.6 * base + 0.4 * maxv * (0.5 + 0.5 * np.sin(xx / 23.0) * np.cos(yy / 17.0))
    sq1 = rng.integers(0, maxv + 1, (64, 64)).astype(np.float64)
    sq2 = rng.integers(0, maxv + 1, (40, 96)).astype(np.float64)
    dt = np.uint8 if a.bits == 8 else np.dtype("<u2")
    with open(a.out, "wb") as f:
        for t in range(a.frames):
            ox, oy = pad + 2 * t, pad + t                      # global pan (+2, +1) px per frame
            y = base[oy:oy + H, ox:ox + W].copy()
            x1, y1 = 40 + 5 * t, 30 + 3 * t                    # rectangle 1: (+5, +3)
            paste(y, sq1, x1, y1)
            x2, y2 = W - 140 - 7 * t, H - 70 - 2 * t           # rectangle 2: (-7, -2)
            paste(y, sq2, x2, y2)
            y = np.clip(np.rint(y + rng.normal(0, 0.004 * maxv * 3, y.shape)), 0, maxv)
            u = np.clip(np.rint(maxv / 2 + 0.1 * (y[::2, ::2] - maxv / 2)), 0, maxv)
            v = np.clip(np.rint(maxv / 2 - 0.1 * (y[::2, ::2] - maxv / 2)), 0, maxv)
            for p in (y, u, v):
                f.write(p.astype(dt).tobytes())


if __name__ == "__main__":
    main()
