// Device helpers shared by the motion-search kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vtmme {

// Device-resident picture: int16 plane with replicated border.  origin = sample (0,0), 16-byte aligned,
// stride (in samples) a multiple of 64.
struct DevPic
{
  int16_t* base;
  int16_t* origin;
  int      stride;
  int      width, height, margin;
};

// ---- MV rate (RdCost.h:301-315) -----------------------------------------------------------------
__device__ __forceinline__ uint32_t eg_bits(int v)
{
  uint32_t len = 1, t = (v <= 0) ? ((uint32_t) (-v) << 1) + 1u : (uint32_t) (v << 1);
  while (t > 128u)
  {
    len += 14;
    t >>= 7;
  }
  return len + ((31u - (uint32_t) __clz(t)) << 1);
}

// getBitsOfVectorWithPredictor with x,y already scaled to quarter-pel (x << costScale).
__device__ __forceinline__ uint32_t mv_bits_q(int xq, int yq, int predQx, int predQy, int imvShift)
{
  return eg_bits((xq - predQx) >> imvShift) + eg_bits((yq - predQy) >> imvShift);
}

// RdCost::getCost: Distortion(m_motionLambda * bits) — one IEEE double multiply, truncated (RdCost.h:191).
__device__ __forceinline__ uint32_t mv_cost(double lambdaMotion, uint32_t bits)
{
  return (uint32_t) (unsigned long long) __dmul_rn(lambdaMotion, (double) bits);
}

// ---- search window (InterSearch::xSetSearchRange, InterSearch.cpp:3496-3535) ----------------------
struct Window
{
  int l, r, t, b;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(hi, max(lo, v)); }

__device__ __forceinline__ int div_pow2_round(int v, int s)   // Mv::divideByPowerOf2, Mv.h:128-136
{
  return (v + (1 << (s - 1)) - (v >= 0 ? 1 : 0)) >> s;
}

// pred in quarter-pel (the frame API's predictor precision); internal precision is 1/16.
__device__ __forceinline__ Window search_window(int predQx, int predQy, int posX, int posY, int picW, int picH,
                                                int ctu, int sr)
{
  const int horMax = (picW + 8 - posX - 1) * 16, horMin = (-ctu - 8 - posX + 1) * 16;
  const int verMax = (picH + 8 - posY - 1) * 16, verMin = (-ctu - 8 - posY + 1) * 16;
  const int px = clampi(predQx * 4, horMin, horMax), py = clampi(predQy * 4, verMin, verMax);
  Window    w;
  w.l = div_pow2_round(clampi(px - sr * 16, horMin, horMax), 4);
  w.r = div_pow2_round(clampi(px + sr * 16, horMin, horMax), 4);
  w.t = div_pow2_round(clampi(py - sr * 16, verMin, verMax), 4);
  w.b = div_pow2_round(clampi(py + sr * 16, verMin, verMax), 4);
  return w;
}

// ---- argmin key: (cost, raster position) lexicographic == xPatternSearch's first strict minimum ----
__device__ __forceinline__ unsigned long long make_key(uint32_t cost, int dx, int dy)
{
  return ((unsigned long long) cost << 32) | ((uint32_t) (dy + 0x8000) << 16) | (uint32_t) (dx + 0x8000);
}
__device__ __forceinline__ int      key_dx(unsigned long long k) { return (int) (k & 0xffffu) - 0x8000; }
__device__ __forceinline__ int      key_dy(unsigned long long k) { return (int) ((k >> 16) & 0xffffu) - 0x8000; }
__device__ __forceinline__ uint32_t key_cost(unsigned long long k) { return (uint32_t) (k >> 32); }

// ---- luma / chroma interpolation coefficients (InterpolationFilter.cpp:57-95,181-216) --------------
static __constant__ int16_t c_lumaFilter[16][8] = {
  { 0, 0, 0, 64, 0, 0, 0, 0 },       { 0, 1, -3, 63, 4, -2, 1, 0 },     { -1, 2, -5, 62, 8, -3, 1, 0 },
  { -1, 3, -8, 60, 13, -4, 1, 0 },   { -1, 4, -10, 58, 17, -5, 1, 0 },  { -1, 4, -11, 52, 26, -8, 3, -1 },
  { -1, 3, -9, 47, 31, -10, 4, -1 }, { -1, 4, -11, 45, 34, -10, 4, -1 }, { -1, 4, -11, 40, 40, -11, 4, -1 },
  { -1, 4, -10, 34, 45, -11, 4, -1 }, { -1, 4, -10, 31, 47, -9, 3, -1 }, { -1, 3, -8, 26, 52, -11, 4, -1 },
  { 0, 1, -5, 17, 58, -10, 4, -1 },  { 0, 1, -4, 13, 60, -8, 3, -1 },   { 0, 1, -3, 8, 62, -5, 2, -1 },
  { 0, 1, -2, 4, 63, -3, 1, 0 }
};
static __constant__ int16_t c_lumaFilter4x4[16][8] = {
  { 0, 0, 0, 64, 0, 0, 0, 0 },     { 0, 1, -3, 63, 4, -2, 1, 0 },   { 0, 1, -5, 62, 8, -3, 1, 0 },
  { 0, 2, -8, 60, 13, -4, 1, 0 },  { 0, 3, -10, 58, 17, -5, 1, 0 }, { 0, 3, -11, 52, 26, -8, 2, 0 },
  { 0, 2, -9, 47, 31, -10, 3, 0 }, { 0, 3, -11, 45, 34, -10, 3, 0 }, { 0, 3, -11, 40, 40, -11, 3, 0 },
  { 0, 3, -10, 34, 45, -11, 3, 0 }, { 0, 3, -10, 31, 47, -9, 2, 0 }, { 0, 2, -8, 26, 52, -11, 3, 0 },
  { 0, 1, -5, 17, 58, -10, 3, 0 }, { 0, 1, -4, 13, 60, -8, 2, 0 },  { 0, 1, -3, 8, 62, -5, 1, 0 },
  { 0, 1, -2, 4, 63, -3, 1, 0 }
};
static __constant__ int16_t c_lumaAltHpel[8]    = { 0, 3, 9, 20, 20, 9, 3, 0 };
static __constant__ int16_t c_chromaFilter[32][4] = {
  { 0, 64, 0, 0 },    { -1, 63, 2, 0 },   { -2, 62, 4, 0 },   { -2, 60, 7, -1 },  { -2, 58, 10, -2 }, { -3, 57, 12, -2 },
  { -4, 56, 14, -2 }, { -4, 55, 15, -2 }, { -4, 54, 16, -2 }, { -5, 53, 18, -2 }, { -6, 52, 20, -2 }, { -6, 49, 24, -3 },
  { -6, 46, 28, -4 }, { -5, 44, 29, -4 }, { -4, 42, 30, -4 }, { -4, 39, 33, -4 }, { -4, 36, 36, -4 }, { -4, 33, 39, -4 },
  { -4, 30, 42, -4 }, { -4, 29, 44, -5 }, { -4, 28, 46, -6 }, { -3, 24, 49, -6 }, { -2, 20, 52, -6 }, { -2, 18, 53, -5 },
  { -2, 16, 54, -4 }, { -2, 15, 55, -4 }, { -2, 14, 56, -4 }, { -2, 12, 57, -3 }, { -2, 10, 58, -2 }, { -1, 7, 60, -2 },
  { 0, 4, 62, -2 },   { 0, 2, 63, -1 }
};

// ---- Hadamard SATD of one TW x TH tile held one row per lane (TH consecutive lanes, aligned to TH) ---
// d[i] = org - cur of column i of this lane's row.  Horizontal butterflies in registers, vertical
// butterflies with warp shuffles, DC term >>2 (JVET_R0164, RdCost.cpp:2258-2261), per-tile normalisation
// of RdCost.cpp:2262/2363/2513/2731.  Returns the tile value in every lane of the tile.
template <int TW, int TH>
__device__ __forceinline__ uint32_t satd_tile_rows(int (&d)[TW], int laneInTile)
{
#pragma unroll
  for (int len = 1; len < TW; len <<= 1)
#pragma unroll
    for (int i = 0; i < TW; i += len << 1)
#pragma unroll
      for (int k = i; k < i + len; k++)
      {
        const int a = d[k], b = d[k + len];
        d[k]       = a + b;
        d[k + len] = a - b;
      }
#pragma unroll
  for (int m = 1; m < TH; m <<= 1)
  {
    const bool up = (laneInTile & m) != 0;
#pragma unroll
    for (int i = 0; i < TW; i++)
    {
      const int o = __shfl_xor_sync(0xffffffffu, d[i], m);
      d[i]        = up ? o - d[i] : d[i] + o;
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 1; i < TW; i++) s += (uint32_t) abs(d[i]);
  const uint32_t dc = (uint32_t) abs(d[0]);
  s += (laneInTile == 0) ? (dc >> 2) : dc;
#pragma unroll
  for (int m = 1; m < TH; m <<= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (TW == 4 && TH == 4) return (s + 1) >> 1;
  if (TW == 8 && TH == 8) return (s + 2) >> 2;
  if (TW * TH == 128)   // (int)(sad / sqrt(16.0*8) * 2): IEEE double divide then multiply, truncate
    return (uint32_t) (int) __dmul_rn(__ddiv_rn((double) (int) s, 0x1.6a09e667f3bcdp+3), 2.0);
  if (TW * TH == 32)    // (int)(sad / sqrt(4.0*8) * 2)
    return (uint32_t) (int) __dmul_rn(__ddiv_rn((double) (int) s, 0x1.6a09e667f3bcdp+2), 2.0);
  return s;   // 2x2
}

// Tile shape VTM picks for a w x h block (RdCost.cpp:2837-2926): returns (tw,th)
__host__ __device__ __forceinline__ void satd_tiling(int w, int h, int& tw, int& th)
{
  if (w > h && (h & 7) == 0 && (w & 15) == 0) { tw = 16; th = 8; }
  else if (w < h && (w & 7) == 0 && (h & 15) == 0) { tw = 8; th = 16; }
  else if (w > h && (h & 3) == 0 && (w & 7) == 0) { tw = 8; th = 4; }
  else if (w < h && (w & 3) == 0 && (h & 7) == 0) { tw = 4; th = 8; }
  else if ((h & 7) == 0 && (w & 7) == 0) { tw = 8; th = 8; }
  else if ((h & 3) == 0 && (w & 3) == 0) { tw = 4; th = 4; }
  else { tw = 2; th = 2; }
}

}   // namespace vtmme
