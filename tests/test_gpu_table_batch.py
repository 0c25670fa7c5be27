"""GPU: the batched table-level kernels on device-resident blocks.
SATD (vtmme_dist_batch kind 1): every block of every shape with min(w, h) >= 8 (the thread-per-tile kernel: 8x8, 16x8 and 8x16
tilings) and a few others (warp-per-block kernel), 10-bit data and the full signed 16-bit range (the kernel is all-integer:
exact for any input), batch sizes that leave partial warps.
Interpolation (vtmme_interp_batch) against the oracle's
InterpolationFilter::filter restatement — many blocks per launch, source rows at even and odd sample addresses (odd row
stride: the parity alternates from row to row), destination rows word-aligned and not, every (isFirst, isLast) pair, 8 and
10 bit.  Exercises the two-outputs-per-thread kernel (interp_pairs_kernel) and, for the shapes it does not take (odd width,
copy), the generic one."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bindings as B  # noqa: E402


@pytest.fixture(scope="module")
def ms():
    import vtm_b200
    m = vtm_b200.MotionSearch(0)
    yield m
    m.close()


SHAPES = [(4, 4), (4, 16), (8, 4), (8, 8), (12, 16), (16, 16), (32, 8), (64, 64), (128, 128), (2, 8), (9, 8)]


@pytest.mark.parametrize("comp", [0, 1])
def test_interp_batch_matches_oracle(ms, oracle_lib, comp):
    import torch
    rng = np.random.default_rng(4100 + comp)
    stream = torch.cuda.Stream()
    ms.set_stream(stream.cuda_stream)
    taps = 8 if comp == 0 else 4
    fracs = (1, 8, 15) if comp == 0 else (1, 16, 31)
    checked = 0
    for (w, h) in SHAPES:
        n = 3 if w * h >= 4096 else 37
        for bd in (10, 8):
            for pad, xoff in ((8, 4), (9, 5)):            # even / odd row stride, even / odd first sample
                ss, sh = w + pad, h + 8
                src = rng.integers(0, 1 << bd, (n, sh, ss), dtype=np.int16)
                mid = rng.integers(-8192, 8192, (n, sh, ss), dtype=np.int16)
                off = 4 * ss + xoff
                for dpad in (0, 1):                       # destination stride w / w + 1 (rows not word-aligned)
                    ds = w + dpad
                    for (vert, first, last) in [(0, 1, 0), (0, 1, 1), (1, 1, 0), (1, 1, 1), (1, 0, 1), (1, 0, 0)]:
                        s = src if first else mid
                        frac = fracs[checked % len(fracs)]
                        alt = 1 if (comp == 0 and frac == 8 and (checked & 1)) else 0
                        checked += 1
                        d_src = torch.from_numpy(s).cuda()
                        d_dst = torch.full((n, h, ds), -1, dtype=torch.int16, device="cuda")
                        with torch.cuda.stream(stream):
                            ms.interp_batch(comp, vert, d_src.data_ptr() + 2 * off, ss, ss * sh, d_dst.data_ptr(), ds, ds * h, w, h,
                                            frac, first, last, bd, alt, n)
                        stream.synchronize()
                        got = d_dst.cpu().numpy()
                        want = np.full((n, h, ds), -1, np.int16)
                        for i in range(n):
                            blk = np.zeros((h, w), np.int16)
                            if vert:
                                oracle_lib.vo_filter_ver(comp, B.ptr(s[i], off), ss, B.ptr(blk), w, w, h, frac, first, last, bd, alt)
                            else:
                                oracle_lib.vo_filter_hor(comp, B.ptr(s[i], off), ss, B.ptr(blk), w, w, h, frac, last, bd, alt)
                            want[i, :, :w] = blk
                        assert np.array_equal(got, want), (comp, w, h, bd, pad, dpad, vert, first, last, frac, alt, taps)
    ms.set_stream(0)
    assert checked > 500


def test_satd_batch_every_block(ms, oracle_lib):
    import torch
    rng = np.random.default_rng(4200)
    sizes = [8, 16, 32, 64, 128]
    shapes = [(w, h) for w in sizes for h in sizes] + [(4, 8), (8, 4), (16, 4), (4, 4), (4, 32)]
    for (w, h) in shapes:
        n = 5 if w * h >= 8192 else (37 if w * h >= 1024 else 301)
        for lo, hi in ((0, 1024), (-32768, 32768)):
            org = rng.integers(lo, hi, (n, h, w)).astype(np.int16)
            cur = rng.integers(lo, hi, (n, h, w)).astype(np.int16)
            d_org, d_cur = torch.from_numpy(org).cuda(), torch.from_numpy(cur).cuda()
            out = torch.full((n,), -1, dtype=torch.int64, device="cuda")
            ms.dist_batch(1, d_org.data_ptr(), w, w * h, d_cur.data_ptr(), w, w * h, w, h, 0, n, out.data_ptr())
            ms.synchronize()
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            want = np.array([oracle_lib.vo_satd(B.ptr(org[i]), w, B.ptr(cur[i]), w, w, h) for i in range(n)], dtype=np.int64)
            assert np.array_equal(got, want), (w, h, lo, hi, np.flatnonzero(got != want)[:5])


def test_interp_tight_buffers(ms, oracle_lib):
    """The source buffer begins with the first sample a filter needs and ends with the last one (the kernels read aligned
    words / 16-byte chunks around them: whatever shares a chunk with a needed sample must meet a zero coefficient)."""
    import torch
    rng = np.random.default_rng(4300)
    stream = torch.cuda.Stream()
    ms.set_stream(stream.cuda_stream)
    for comp, taps in ((0, 8), (1, 4)):
        before = taps // 2 - 1
        for (w, h) in [(8, 8), (16, 8), (64, 16), (128, 128), (4, 8)]:
            for vert in (0, 1):
                for lead in (0, 1, 3):      # samples in front of the tight buffer inside the allocation: shifts its alignment
                    ss = w + (0 if vert else taps - 1)
                    rows = h + (taps - 1 if vert else 0)
                    flat = rng.integers(0, 1024, lead + rows * ss, dtype=np.int16)
                    tight = flat[lead:]
                    off = before * ss if vert else before
                    frac = 5 if comp == 0 else 11
                    want = np.zeros((h, w), np.int16)
                    if vert:
                        oracle_lib.vo_filter_ver(comp, B.ptr(tight, off), ss, B.ptr(want), w, w, h, frac, 1, 1, 10, 0)
                    else:
                        oracle_lib.vo_filter_hor(comp, B.ptr(tight, off), ss, B.ptr(want), w, w, h, frac, 1, 10, 0)
                    d_src = torch.from_numpy(flat).cuda()
                    d_dst = torch.zeros((h, w), dtype=torch.int16, device="cuda")
                    with torch.cuda.stream(stream):
                        ms.interp_batch(comp, vert, d_src.data_ptr() + 2 * (lead + off), ss, ss * rows, d_dst.data_ptr(), w, w * h, w, h,
                                        frac, 1, 1, 10, 0, 1)
                    stream.synchronize()
                    assert np.array_equal(d_dst.cpu().numpy(), want), (comp, w, h, vert, lead)
    ms.set_stream(0)
