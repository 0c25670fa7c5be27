// Integer-pel / 4-pel AMVR refinement, InterSearch::xPatternSearchIntRefine (EncoderLib/InterSearch.cpp:4172-4282):
// after the integer search of a CU coded with imv = IMV_FPEL / IMV_4PEL, the best position and its eight
// neighbours in AMVR units are tried against both AMVP candidates (SATD when HadamardME, else SAD, weighted by the
// bi-prediction weight), the rate being the MVD bits in AMVR units plus the MVP index bits.
//
// The (position, candidate) probes that need a distortion of their own are spread, tile by tile, over all lanes
// of the CTA (or of several CTAs for patterns larger than 32x32); the decision replays the reference's loop.
#pragma once
#include "me_common.cuh"
#include "me_kernels.h"

namespace vtmme {

constexpr int kIntRefProbes = 18;   // 9 positions x 2 AMVP candidates, probe = pos * 2 + cand

struct IntRefSmem
{
  int      off[kIntRefProbes];    // element offset of the probe's block from refAtPU
  int      list[kIntRefProbes];   // probes with a distortion of their own
  int      nList;
  uint32_t acc[kIntRefProbes];
};

static __constant__ int8_t c_intRefinePos[9][2] = { { 0, 0 },  { -1, -1 }, { -1, 0 }, { -1, 1 }, { 0, -1 },
                                                    { 0, 1 },  { 1, -1 },  { 1, 0 },  { 1, 1 } };   // InterSearch.cpp:4195

// Mv::changePrecision (CommonLib/Mv.h:183-197) by `shift` bits: left shift, or right shift to nearest, ties toward zero
__device__ __forceinline__ int change_prec(int v, int shift)
{
  if (shift >= 0) return v * (1 << shift);
  const int rs = -shift, off = 1 << (rs - 1);
  return v >= 0 ? (v + off - 1) >> rs : (v + off) >> rs;
}

// bits between MV_PRECISION_INTERNAL and the AMVR precision of cu.imv (Mv::m_amvrPrecision, Mv.cpp:41)
__device__ __forceinline__ int amvr_shift(int imv) { return imv == 1 ? 4 : (imv == 2 ? 6 : (imv == 3 ? 3 : 2)); }

// cTestMv[cand] of position `pos` (:4214-4217); rcMv = (mvX16, mvY16)
__device__ __forceinline__ void intrefine_test_mv(const DevAmvr& a, int mvX16, int mvY16, int pos, int cand, int& tx, int& ty)
{
  const int s  = amvr_shift(a.imv);
  const int bx = change_prec(change_prec(mvX16 - a.candX[cand], -s), s);   // roundTransPrecInternal2Amvr, :4203-4204
  const int by = change_prec(change_prec(mvY16 - a.candY[cand], -s), s);
  tx           = change_prec(c_intRefinePos[pos][0], s) + bx + a.candX[cand];
  ty           = change_prec(c_intRefinePos[pos][1], s) + by + a.candY[cand];
}

// clipMvInPic (CommonLib/Mv.cpp:53-71) and the sample offset of :4239
__device__ __forceinline__ int intrefine_offset(const DevAmvr& a, int tx, int ty, int refStride)
{
  const int horMax = (a.picW + 8 - a.posX - 1) * 16, horMin = (-a.maxCuW - 8 - a.posX + 1) * 16;
  const int verMax = (a.picH + 8 - a.posY - 1) * 16, verMin = (-a.maxCuH - 8 - a.posY + 1) * 16;
  const int cx = clampi(tx, horMin, horMax), cy = clampi(ty, verMin, verMax);
  return (cy >> 4) * refStride + (cx >> 4);
}

// one TW x TH tile of one probe: TH consecutive lanes, one row each
template <int TW, int TH, bool HAD>
__device__ __forceinline__ uint32_t intrefine_tile(const int16_t* org, int os, const int16_t* cur, int cs, int tx, int ty,
                                                   int laneInTile)
{
  const int16_t* op = org + (ptrdiff_t) (ty + laneInTile) * os + tx;
  const int16_t* cp = cur + (ptrdiff_t) (ty + laneInTile) * cs + tx;
  int            d[TW];
#pragma unroll
  for (int i = 0; i < TW; i++) d[i] = (int) op[i] - (int) cp[i];
  if (HAD) return satd_tile_rows<TW, TH>(d, laneInTile);
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < TW; i++) s += (uint32_t) abs(d[i]);
#pragma unroll
  for (int m = 1; m < TH; m <<= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  return s;
}

template <int TW, int TH, bool HAD, int THREADS>
__device__ __forceinline__ void intrefine_units(IntRefSmem& s, const int16_t* org, int os, const int16_t* refAtPU, int rs,
                                                int w, int h, int part, int nParts)
{
  const int tilesX = w / TW, nTiles = tilesX * (h / TH);
  const int nGroups = THREADS / TH, group = threadIdx.x / TH, laneInTile = threadIdx.x % TH;
  const int units = s.nList * nTiles;
  // all lanes of a warp reach the shuffles together: warp-uniform rounds
  for (int u0 = part * nGroups; u0 < units; u0 += nParts * nGroups)
  {
    const int  u      = u0 + group;
    const bool active = u < units;
    const int  uu     = active ? u : 0;
    const int  k = uu / nTiles, t = uu - k * nTiles;
    const int  p = s.list[k];
    const uint32_t v = intrefine_tile<TW, TH, HAD>(org, os, refAtPU + s.off[p], rs, (t % tilesX) * TW, (t / tilesX) * TH, laneInTile);
    if (active && laneInTile == 0) atomicAdd(&s.acc[p], v);
  }
}

// Distortions of the probes into s.acc[] (raw, before the bi-prediction weight).  Slice `part` of `nParts` of the
// (probe, tile) units; the whole CTA of THREADS threads takes part.  (mvX, mvY) = integer MV found by the search.
template <int THREADS>
__device__ inline void intrefine_accumulate(IntRefSmem& s, const DevJob& j, const int16_t* org, int os, int mvX, int mvY,
                                            int part, int nParts)
{
  const DevAmvr& a = j.amvr;
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid < kIntRefProbes)
  {
    const int pos = tid >> 1, cand = tid & 1;
    int       need = 0, off = 0;
    if (cand < a.numCand)
    {
      int tx, ty;
      intrefine_test_mv(a, mvX * 16, mvY * 16, pos, cand, tx, ty);
      need = 1;
      if (cand == 1)
      {
        int ox, oy;
        intrefine_test_mv(a, mvX * 16, mvY * 16, pos, 0, ox, oy);
        need = ox != tx || oy != ty;   // :4232 — same MV as candidate 0: its distortion is reused
      }
      off = intrefine_offset(a, tx, ty, j.refStride);
    }
    s.off[tid] = off;
    s.acc[tid] = need ? 0u : 0xffffffffu;
  }
  __syncthreads();
  if (tid == 0)
  {
    int n = 0;
    for (int p = 0; p < kIntRefProbes; p++)
      if (s.acc[p] == 0u) s.list[n++] = p;
    s.nList = n;
  }
  __syncthreads();
  int tw, th;
  if (j.useHad)
    satd_tiling(j.w, j.h, tw, th);
  else
  {
    tw = j.w >= 8 ? 8 : 4;
    th = j.h >= 8 ? 8 : 4;
  }
  const int16_t* ref = j.refAtPU;
  const int      rs  = j.refStride;
  if (j.useHad)
  {
    if (tw == 8 && th == 8) intrefine_units<8, 8, true, THREADS>(s, org, os, ref, rs, j.w, j.h, part, nParts);
    else if (tw == 16) intrefine_units<16, 8, true, THREADS>(s, org, os, ref, rs, j.w, j.h, part, nParts);
    else if (th == 16) intrefine_units<8, 16, true, THREADS>(s, org, os, ref, rs, j.w, j.h, part, nParts);
    else if (tw == 8 && th == 4) intrefine_units<8, 4, true, THREADS>(s, org, os, ref, rs, j.w, j.h, part, nParts);
    else if (tw == 4 && th == 8) intrefine_units<4, 8, true, THREADS>(s, org, os, ref, rs, j.w, j.h, part, nParts);
    else intrefine_units<4, 4, true, THREADS>(s, org, os, ref, rs, j.w, j.h, part, nParts);
  }
  else
  {
    if (tw == 8 && th == 8) intrefine_units<8, 8, false, THREADS>(s, org, os, ref, rs, j.w, j.h, part, nParts);
    else if (tw == 8) intrefine_units<8, 4, false, THREADS>(s, org, os, ref, rs, j.w, j.h, part, nParts);
    else if (th == 8) intrefine_units<4, 8, false, THREADS>(s, org, os, ref, rs, j.w, j.h, part, nParts);
    else intrefine_units<4, 4, false, THREADS>(s, org, os, ref, rs, j.w, j.h, part, nParts);
  }
  __syncthreads();
}

// The loop of :4207-4262 and the tail :4263-4280 over the probe distortions dist[pos*2+cand] (entries of reused
// probes are ignored).  Serial, one thread.
__device__ inline void intrefine_decide(const DevJob& j, int mvX, int mvY, const uint32_t* dist, DevJobResult& res)
{
  const DevAmvr& a = j.amvr;
  const int      s = amvr_shift(a.imv);
  unsigned long long best = ~0ull, satd = 0;
  int      bestX = mvX * 16, bestY = mvY * 16, bestIdx = a.mvpIdx;
  uint32_t bestBits = 0;
  uint32_t bits = a.bits - a.mvpIdxBits[a.mvpIdx];   // :4188
  int      t0x = 0, t0y = 0;
  for (int pos = 0; pos < 9; pos++)
    for (int i = 0; i < a.numCand; i++)
    {
      int tx, ty;
      intrefine_test_mv(a, mvX * 16, mvY * 16, pos, i, tx, ty);
      unsigned long long d;
      if (i == 0 || tx != t0x || ty != t0y)
        d = satd = (unsigned long long) __dmul_rn((double) dist[pos * 2 + i], a.fWeight);   // :4240
      else
        d = satd;
      if (i == 0)
      {
        t0x = tx;
        t0y = ty;
      }
      const uint32_t mvBits = a.mvpIdxBits[i] + eg_bits(change_prec(tx, -s) - change_prec(a.candX[i], -s)) +
                              eg_bits(change_prec(ty, -s) - change_prec(a.candY[i], -s));   // :4247-4254, cost scale 0
      d += mv_cost(j.lambda, mvBits);
      if (d < best)
      {
        best     = d;
        bestX    = tx;
        bestY    = ty;
        bestIdx  = i;
        bestBits = mvBits;
      }
    }
  res.amvrMvX = bestX;
  res.amvrMvY = bestY;
  res.mvpIdx  = bestIdx;
  bits += bestBits;
  res.bits = bits;
  res.cost = best - mv_cost(j.lambda, bestBits) + mv_cost(j.lambda, bits);   // :4276
}

}   // namespace vtmme
