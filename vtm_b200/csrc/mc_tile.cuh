// One <=16x16 tile of InterPrediction::xPredInterBlk (CommonLib/InterPrediction.cpp:660-830, plain path) by one warp: shared by
// the motion-compensation kernels (mc_kernels.cu) and the symmetric-MVD search (smvd_kernels.cu).
#pragma once
#include "me_kernels.h"

namespace vtmme {
namespace mc {

constexpr int kTile       = 16;
constexpr int kPatchRows  = kTile + 7;
constexpr int kPatchPitch = kTile + 8;   // 24 int16 = 12 words: rows of a column walk hit distinct banks pairwise

// rounding of one filter stage, InterpolationFilter::filter<> (InterpolationFilter.cpp:578-603)
__device__ __forceinline__ void stage_rounding(bool isFirst, bool isLast, int hr, int& shift, int& offset)
{
  shift = 6;
  if (isLast)
  {
    shift += isFirst ? 0 : hr;
    offset = 1 << (shift - 1);
    offset += isFirst ? 0 : (8192 << 6);
  }
  else
  {
    shift -= isFirst ? hr : 0;
    offset = isFirst ? -(8192 << shift) : 0;
  }
}

template <int TAPS>
__device__ __forceinline__ void load_coeff(int (&c)[TAPS], int frac, bool q4, bool alt)
{
  if (TAPS == 8)
  {
    // coefficient choice of filterHor / filterVer (InterpolationFilter.cpp:782-794, 865-877)
    const int16_t* t = (frac == 8 && alt) ? c_lumaAltHpel : (q4 ? c_lumaFilter4x4[frac] : c_lumaFilter[frac]);
#pragma unroll
    for (int k = 0; k < TAPS; k++) c[k] = t[k];
  }
  else
  {
#pragma unroll
    for (int k = 0; k < TAPS; k++) c[k] = c_chromaFilter[frac][k];
  }
}

// One tile by one warp.  sink(y, x, v) receives every prediction sample of the tile.
template <int TAPS, class Sink>
__device__ __forceinline__ void mc_tile(const McTile& t, int bi, int bitDepth, int alt, int16_t* patch, int16_t* mid, int lane,
                                        Sink& sink)
{
  constexpr int above = TAPS / 2 - 1, halo = TAPS - 1;
  const bool px = t.xFrac != 0, py = t.yFrac != 0;
  const int  tw = t.tw, th = t.th;
  const int  c0 = px ? -above : 0, pw = tw + (px ? halo : 0);
  const int  r0 = py ? -above : 0, ph = th + (py ? halo : 0);

  if (t.winMaxX)
  {
    // the filter reads the prefetched window with replicated borders instead of the picture (InterPrediction.cpp:1710-1730)
    for (int i = lane; i < pw * ph; i += 32)
    {
      const int r = i / pw, c = i - r * pw;
      const int wr = min(max((int) t.winY + r0 + r, 0), (int) t.winMaxY), wc = min(max((int) t.winX + c0 + c, 0), (int) t.winMaxX);
      patch[r * kPatchPitch + c] = t.src[(ptrdiff_t) wr * t.srcStride + wc];
    }
  }
  else
  {
    for (int i = lane; i < pw * ph; i += 32)
    {
      const int r = i / pw, c = i - r * pw;
      patch[r * kPatchPitch + c] = t.src[(ptrdiff_t) (r0 + r) * t.srcStride + c0 + c];
    }
  }
  __syncwarp();

  const int  hr   = max(2, 14 - bitDepth);
  const int  maxv = (1 << bitDepth) - 1;
  const bool rnd  = !bi;   // rndRes (InterPrediction.cpp:673)
  const int  outs = tw * th;

  if (!px && !py)
  {
    // filterCopy<true, rnd> (InterpolationFilter.cpp:397-470)
    for (int o = lane; o < outs; o += 32)
    {
      const int y = o / tw, x = o - y * tw;
      const int v = patch[y * kPatchPitch + x];
      sink(y, x, rnd ? v : (int) (int16_t) ((int16_t) (v << hr) - (int16_t) 8192));
    }
    return;
  }
  int shift, offset;
  if (px)
  {
    int cf[TAPS];
    load_coeff<TAPS>(cf, t.xFrac, t.q4Hor != 0, alt != 0);
    const bool last = rnd && !py;
    stage_rounding(true, last, hr, shift, offset);
    const int n = tw * ph;   // with a vertical pass to follow: every staged row
    for (int o = lane; o < n; o += 32)
    {
      const int      y = o / tw, x = o - y * tw;
      const int16_t* sp = patch + y * kPatchPitch + x;
      int            sum = 0;
#pragma unroll
      for (int k = 0; k < TAPS; k++) sum += (int) sp[k] * cf[k];
      int v = (int16_t) ((sum + offset) >> shift);
      if (!py)
      {
        if (last) v = min(max(v, 0), maxv);
        sink(y, x, v);
      }
      else
        mid[y * kTile + x] = (int16_t) v;
    }
    if (!py) return;
    __syncwarp();
  }
  {
    int cf[TAPS];
    load_coeff<TAPS>(cf, t.yFrac, t.q4Ver != 0, alt != 0);
    stage_rounding(!px, rnd, hr, shift, offset);
    const int16_t* plane = px ? mid : patch;
    const int      pitch = px ? kTile : kPatchPitch;
    for (int o = lane; o < outs; o += 32)
    {
      const int      y = o / tw, x = o - y * tw;
      const int16_t* sp = plane + y * pitch + x;
      int            sum = 0;
#pragma unroll
      for (int k = 0; k < TAPS; k++) sum += (int) sp[k * pitch] * cf[k];
      int v = (int16_t) ((sum + offset) >> shift);
      if (rnd) v = min(max(v, 0), maxv);
      sink(y, x, v);
    }
  }
}

}   // namespace mc
}   // namespace vtmme
