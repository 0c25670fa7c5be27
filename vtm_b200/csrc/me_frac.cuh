// Fractional-pel refinement, the body of InterSearch::xPatternSearchFracDIF (EncoderLib/InterSearch.cpp:4296-4338)
// with xExtDIFUpSamplingH/Q (:5840-6050) and xPatternRefinement (:707-761) fused: the separable 8-tap
// interpolation (InterpolationFilter.cpp:550-656, 14-bit intermediates) feeds the Hadamard SATD
// (RdCost.cpp:2140-2934) directly, nothing goes back to global memory.
//
// Direct form (the tests check against the CPU restatement that it equals the reference's m_filteredBlock
// planes): the candidate at quarter-pel offset (dqx,dqy) from the integer MV is interpolated at integer
// base (dq >> 2) with phase (dq & 3).  One CTA refines one CU, in chunks of at most 32x32 samples:
//   1. stage the chunk's original samples and the (cw+8)x(ch+8) reference patch in shared memory,
//   2. horizontal pass for the three dqx of the stage into 14-bit planes,
//   3. every (candidate, SATD tile) pair: vertical pass on the fly, diff, Hadamard butterflies
//      (in-register horizontally, warp shuffles vertically), accumulate per candidate.
#pragma once
#include "me_common.cuh"

namespace vtmme {

constexpr int kFracThreads = 128;
constexpr int kFracChunk   = 32;
constexpr int kPatchStride = kFracChunk + 8;                      // samples
constexpr int kPatchRows   = kFracChunk + 8;
constexpr int kPlaneStride = kFracChunk;                          // 14-bit plane: [ch+8][cw]
constexpr int kPlaneSize   = kPlaneStride * (kFracChunk + 8);

struct FracSmem
{
  int16_t  org[kFracChunk * kFracChunk];
  uint16_t patch[kPatchRows * kPatchStride];
  int16_t  plane[3][kPlaneSize];
  uint32_t acc[9];
  uint32_t centre;   // raw distortion of candidate 0 of the last stage
  int      best;
};

struct FracJob
{
  const int16_t* org;       // pattern, global memory
  int            orgStride;
  const int16_t* refAtMv;   // reference plane at PU position + integer MV
  int            refStride;
  int            w, h;
  int            mvX, mvY;  // integer MV (for the rate term)
  int            predQx, predQy;
  int            bitDepth, useHad, useAltHpel, imvShift;
  double         lambda;
};

struct FracOut
{
  int      halfX, halfY, qterX, qterY;
  uint32_t cost;
};

static __constant__ int8_t c_refineH[9][2] = { { 0, 0 },  { 0, -1 }, { 0, 1 },  { -1, 0 }, { 1, 0 },
                                                   { -1, -1 }, { 1, -1 }, { -1, 1 }, { 1, 1 } };   // InterSearch.cpp:59-70
static __constant__ int8_t c_refineQ[9][2] = { { 0, 0 },  { 0, -1 }, { 0, 1 },  { -1, -1 }, { 1, -1 },
                                                   { -1, 0 }, { 1, 0 },  { -1, 1 }, { 1, 1 } };   // :72-83

// distortion of one (candidate, tile): TW x TH tile at (tx,ty) of the chunk, candidate vertical offset dqy,
// horizontal plane pl.  Executed by TH consecutive lanes; returns the tile value in all of them.
template <int TW, int TH, bool HAD>
__device__ __forceinline__ uint32_t frac_tile(const FracSmem& sm, int pl, int dqy, int tx, int ty, int laneInTile,
                                              int bitDepth, bool altV)
{
  const int iy = dqy >> 2, py = dqy & 3;
  const int y  = ty + laneInTile;
  const int hr = max(2, 14 - bitDepth);
  const int maxv = (1 << bitDepth) - 1;
  int       d[TW];
  const int16_t* pp = sm.plane[pl] + (y + iy + 4) * kPlaneStride + tx;   // plane row r holds picture row r-4
  if (py == 0)
  {
#pragma unroll
    for (int i = 0; i < TW; i++)
    {
      int v = (int16_t) ((pp[i] + 8192 + (1 << (hr - 1))) >> hr);   // filterCopy<false,true>, InterpolationFilter.cpp:508-520
      d[i]  = min(max(v, 0), maxv);
    }
  }
  else
  {
    const int16_t* c = altV ? c_lumaAltHpel : c_lumaFilter[py * 4];
    int            cf[8];
#pragma unroll
    for (int k = 0; k < 8; k++) cf[k] = c[k];
    const int shift  = 6 + hr;
    const int offset = (1 << (shift - 1)) + (8192 << 6);
    int       sum[TW];
#pragma unroll
    for (int i = 0; i < TW; i++) sum[i] = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
      const int16_t* row = pp + (k - 3) * kPlaneStride;
#pragma unroll
      for (int i = 0; i < TW; i++) sum[i] += row[i] * cf[k];
    }
#pragma unroll
    for (int i = 0; i < TW; i++)
    {
      int v = (int16_t) ((sum[i] + offset) >> shift);   // filter<8,true,false,true>, InterpolationFilter.cpp:592-650
      d[i]  = min(max(v, 0), maxv);
    }
  }
  const int16_t* op = sm.org + y * kFracChunk + tx;
#pragma unroll
  for (int i = 0; i < TW; i++) d[i] = op[i] - d[i];
  if (HAD) return satd_tile_rows<TW, TH>(d, laneInTile);
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < TW; i++) s += (uint32_t) abs(d[i]);
#pragma unroll
  for (int m = 1; m < TH; m <<= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  return s;
}

template <int TW, int TH, bool HAD>
__device__ __forceinline__ void frac_chunk_tiles(FracSmem& sm, int cw, int ch, const int (&dqy)[9], const int (&plane)[9],
                                                 int bitDepth, bool altV)
{
  const int tilesX = cw / TW, tilesY = ch / TH, nTiles = tilesX * tilesY;
  const int group = threadIdx.x / TH, laneInTile = threadIdx.x % TH, nGroups = kFracThreads / TH;
  const int units = 9 * nTiles;
  // all lanes of a warp must reach the shuffles together: iterate in warp-uniform rounds
  const int rounds = (units + nGroups - 1) / nGroups;
  for (int it = 0; it < rounds; it++)
  {
    const int  u      = it * nGroups + group;
    const bool active = u < units;
    const int  uu     = active ? u : 0;
    const int  c = uu / nTiles, t = uu - c * nTiles;
    const int  tx = (t % tilesX) * TW, ty = (t / tilesX) * TH;
    const uint32_t v = frac_tile<TW, TH, HAD>(sm, plane[c], dqy[c], tx, ty, laneInTile, bitDepth, altV);
    if (active && laneInTile == 0) atomicAdd(&sm.acc[c], v);
  }
}

// xPatternRefinement's choice (InterSearch.cpp:727-756) from the nine candidate distortions of a stage around
// (cqx,cqy): first strict minimum in list order.  Serial, one thread.
__device__ inline void frac_pick(const uint32_t* dist, const FracJob& j, int cqx, int cqy, int step, int& bestDir,
                                 uint32_t& bestCost)
{
  const int8_t (*tab)[2] = (step == 2) ? c_refineH : c_refineQ;
  bestCost = 0xffffffffu;
  bestDir  = 0;
  for (int i = 0; i < 9; i++)
  {
    const uint32_t cost = dist[i] + mv_cost(j.lambda, mv_bits_q(j.mvX * 4 + cqx + tab[i][0] * step, j.mvY * 4 + cqy + tab[i][1] * step,
                                                                j.predQx, j.predQy, 0));
    if (cost < bestCost)
    {
      bestCost = cost;
      bestDir  = i;
    }
  }
}

__device__ __forceinline__ int frac_num_chunks(int w, int h)
{
  const int cw = min(w, kFracChunk), ch = min(h, kFracChunk);
  return (w / cw) * (h / ch);
}

// Candidate distortions of one refinement stage around (cqx,cqy) (quarter-pel offset from the integer MV) with step
// `step` (2 = half-pel, 1 = quarter-pel), accumulated into sm.acc[0..8] over chunk `chunkSel` (or all chunks if < 0).
// `staged`: the (single) chunk's original samples and reference patch are already in sm from the previous stage.
__device__ inline void frac_stage_sums(FracSmem& sm, const FracJob& j, int cqx, int cqy, int step, bool alt, int chunkSel,
                                       bool staged = false)
{
  const int8_t (*tab)[2] = (step == 2) ? c_refineH : c_refineQ;
  const int tid = threadIdx.x;
  if (tid < 9) sm.acc[tid] = 0;
  int dqy[9], plane[9];
#pragma unroll
  for (int i = 0; i < 9; i++)
  {
    dqy[i]   = cqy + tab[i][1] * step;
    plane[i] = tab[i][0] + 1;   // plane 0,1,2 = dqx of cqx-step, cqx, cqx+step
  }
  const int hr = max(2, 14 - j.bitDepth);
  int       tw, th;
  if (j.useHad)
    satd_tiling(j.w, j.h, tw, th);
  else
  {
    tw = j.w >= 8 ? 8 : 4;
    th = 8;
    if (j.h < 8) th = 4;
  }
  const int cw = min(j.w, kFracChunk), ch = min(j.h, kFracChunk);

  const int chunksX = j.w / cw;
  int       chunk   = 0;
  for (int cy0 = 0; cy0 < j.h; cy0 += ch)
    for (int cx0 = 0; cx0 < j.w; cx0 += cw, chunk++)
    {
      if (chunkSel >= 0 && chunk != chunkSel) continue;
      __syncthreads();
      // 1. stage original chunk and reference patch rows [-4, ch+4), cols [-4, cw+4)
      if (!staged)
      {
        for (int i = tid; i < cw * ch; i += kFracThreads)
        {
          const int y = i / cw, x = i - y * cw;
          sm.org[y * kFracChunk + x] = j.org[(size_t) (cy0 + y) * j.orgStride + cx0 + x];
        }
        for (int i = tid; i < (ch + 8) * (cw + 8); i += kFracThreads)
        {
          const int y = i / (cw + 8), x = i - y * (cw + 8);
          sm.patch[y * kPatchStride + x] =
            (uint16_t) j.refAtMv[(ptrdiff_t) (cy0 + y - 4) * j.refStride + (cx0 + x - 4)];
        }
        __syncthreads();
      }
      // 2. horizontal pass: plane[p][r][c], r in [0,ch+8) <-> picture row r-4, for dqx = cqx + (p-1)*step
      for (int i = tid; i < 3 * (ch + 8) * cw; i += kFracThreads)
      {
        const int p = i / ((ch + 8) * cw), rem = i - p * (ch + 8) * cw;
        const int r = rem / cw, c = rem - r * cw;
        const int dq = cqx + (p - 1) * step;
        const int ix = dq >> 2, px = dq & 3;
        const uint16_t* src = sm.patch + r * kPatchStride + (c + ix + 4);
        int             v;
        if (px == 0)
          v = (int16_t) ((int16_t) (src[0] << hr) - (int16_t) 8192);   // filterCopy<true,false>
        else
        {
          const int16_t* cf = (alt && px == 2) ? c_lumaAltHpel : c_lumaFilter[px * 4];
          int            sum = 0;
#pragma unroll
          for (int k = 0; k < 8; k++) sum += (int) src[k - 3] * cf[k];
          const int shift = 6 - hr;
          v               = (int16_t) ((sum - (8192 << shift)) >> shift);   // filter<8,false,true,false>
        }
        sm.plane[p][r * kPlaneStride + c] = (int16_t) v;
      }
      __syncthreads();
      // 3. candidates x tiles
      const bool altV = alt;
      if (j.useHad)
      {
        if (tw == 8 && th == 8) frac_chunk_tiles<8, 8, true>(sm, cw, ch, dqy, plane, j.bitDepth, altV);
        else if (tw == 16) frac_chunk_tiles<16, 8, true>(sm, cw, ch, dqy, plane, j.bitDepth, altV);
        else if (th == 16) frac_chunk_tiles<8, 16, true>(sm, cw, ch, dqy, plane, j.bitDepth, altV);
        else if (tw == 8 && th == 4) frac_chunk_tiles<8, 4, true>(sm, cw, ch, dqy, plane, j.bitDepth, altV);
        else if (tw == 4 && th == 8) frac_chunk_tiles<4, 8, true>(sm, cw, ch, dqy, plane, j.bitDepth, altV);
        else frac_chunk_tiles<4, 4, true>(sm, cw, ch, dqy, plane, j.bitDepth, altV);
      }
      else
      {
        if (tw == 8 && th == 8) frac_chunk_tiles<8, 8, false>(sm, cw, ch, dqy, plane, j.bitDepth, altV);
        else if (tw == 8) frac_chunk_tiles<8, 4, false>(sm, cw, ch, dqy, plane, j.bitDepth, altV);
        else if (th == 8) frac_chunk_tiles<4, 8, false>(sm, cw, ch, dqy, plane, j.bitDepth, altV);
        else frac_chunk_tiles<4, 4, false>(sm, cw, ch, dqy, plane, j.bitDepth, altV);
      }
    }
  (void) chunksX;
  __syncthreads();
}

// One whole stage (all chunks) and its decision.  Returns the winning direction index in sm.best and its cost.
__device__ inline uint32_t frac_stage(FracSmem& sm, const FracJob& j, int cqx, int cqy, int step, bool alt, bool staged = false)
{
  frac_stage_sums(sm, j, cqx, cqy, step, alt, -1, staged);
  if (threadIdx.x < 32)
  {
    // frac_pick's choice with one candidate per lane: min over (cost, list index) = first strict minimum in list order
    const int8_t (*tab)[2] = (step == 2) ? c_refineH : c_refineQ;
    const int          i   = threadIdx.x;
    unsigned long long key = ~0ull;
    uint32_t           d0  = 0;
    if (i < 9)
    {
      const uint32_t dist = sm.acc[i];
      d0                  = dist;
      const uint32_t cost = dist + mv_cost(j.lambda, mv_bits_q(j.mvX * 4 + cqx + tab[i][0] * step, j.mvY * 4 + cqy + tab[i][1] * step,
                                                               j.predQx, j.predQy, 0));
      key = ((unsigned long long) cost << 8) | (unsigned) i;
    }
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1)
    {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, m);
      key                        = o < key ? o : key;
    }
    __syncwarp();
    if (i == 0)
    {
      sm.best   = (int) (key & 0xffu);
      sm.centre = d0;
      sm.acc[0] = (uint32_t) (key >> 8);
    }
  }
  __syncthreads();
  const uint32_t cost = sm.acc[0];
  __syncthreads();
  return cost;
}

// Whole xPatternSearchFracDIF body for one CU, executed by a CTA of kFracThreads threads.
__device__ inline FracOut frac_refine_cta(FracSmem& sm, const FracJob& j)
{
  FracOut o;
  o.halfX = o.halfY = o.qterX = o.qterY = 0;
  o.cost = frac_stage(sm, j, 0, 0, 2, j.useAltHpel != 0);
  const int hd = sm.best;
  o.halfX = c_refineH[hd][0];
  o.halfY = c_refineH[hd][1];
  if (j.imvShift == 0)
  {
    o.cost = frac_stage(sm, j, o.halfX * 2, o.halfY * 2, 1, false, frac_num_chunks(j.w, j.h) == 1);
    const int qd = sm.best;
    o.qterX = c_refineQ[qd][0];
    o.qterY = c_refineQ[qd][1];
  }
  return o;
}

}   // namespace vtmme
