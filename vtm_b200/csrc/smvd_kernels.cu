// Symmetric-MVD search: InterSearch::xSymmetricMotionEstimation (EncoderLib/InterSearch.cpp:4506-4518) with
// xSymmeticRefineMvSearch (:4393-4503) and xGetSymmetricCost (:4341-4391).
//
// The search moves the MV of the searched list on a diamond (at most 8 >> imv rounds) and then once on a cross; the MV of the
// other list mirrors the MV difference.  The candidates of a round are fixed by the centre the round starts from, the
// reference tests them one after the other with a strict `<`: so a round is evaluated in parallel — every (candidate,
// 16x16 tile) is one unit of work for one warp — and one thread replays the sequential acceptance over the candidates'
// costs.  A unit is: both 8-tap predictions of the tile at the clipped MVs (the tile routine of vtmme_mc_batch, uni-
// directional rounding), 2*org - predA (removeHighFreq; removeWeightHighFreq under a BCW weight), SATD in the reference's
// tiling of the CU (8x8, 16x8 or 8x16 tiles, RdCost.cpp:2837-2926) or SAD against predB.  One CTA per search; a launch
// takes any number of searches.
#include "mc_tile.cuh"
#include "me_kernels.h"

namespace vtmme {
namespace {

using namespace mc;

constexpr int kSmvdWarps = 8;

__device__ __forceinline__ int change_prec(int v, int shift)   // Mv::changePrecision, CommonLib/Mv.h:183-197
{
  if (shift >= 0) return v * (1 << shift);
  const int rs = -shift, off = 1 << (rs - 1);
  return v >= 0 ? (v + off - 1) >> rs : (v + off) >> rs;
}

__device__ __forceinline__ void clip_mv(int& mx, int& my, const DevSmvd& j)   // clipMvInPic, CommonLib/Mv.cpp:53-71
{
  const int horMax = (j.picW + 8 - j.x - 1) << 4, horMin = (-j.maxCu - 8 - j.x + 1) * 16;
  const int verMax = (j.picH + 8 - j.y - 1) << 4, verMin = (-j.maxCu - 8 - j.y + 1) * 16;
  mx = min(horMax, max(horMin, mx));
  my = min(verMax, max(verMin, my));
}

struct TileSink
{
  int16_t* dst;   // [16][16]
  __device__ __forceinline__ void operator()(int y, int x, int v) const { dst[y * kTile + x] = (int16_t) v; }
};

struct SmvdState
{
  int                curX, curY, tarX, tarY;
  unsigned long long cost;
  int                start, end, stop;
};

__global__ void __launch_bounds__(kSmvdWarps * 32) smvd_search_kernel(const DevSmvd* __restrict__ jobs, DevSmvdResult* __restrict__ out)
{
  __shared__ int16_t            s_patch[kSmvdWarps][kPatchRows * kPatchPitch];
  __shared__ int16_t            s_mid[kSmvdWarps][kPatchRows * kTile];
  __shared__ int16_t            s_predA[kSmvdWarps][kTile * kTile];
  __shared__ int16_t            s_predB[kSmvdWarps][kTile * kTile];
  __shared__ unsigned long long s_dist[8];
  __shared__ SmvdState          st;

  const DevSmvd j    = jobs[blockIdx.x];
  const int     warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int     tilesX = (j.w + 15) >> 4, nTiles = tilesX * ((j.h + 15) >> 4);
  const int     maxv = (1 << j.bd) - 1;
  const int     alt  = j.imv == 3;
  // AMVR: Mv::m_amvrPrecision (Mv.cpp:43) and the step of the patterns (:4512)
  const int down      = j.imv == 0 ? -2 : (j.imv == 1 ? -4 : (j.imv == 2 ? -6 : -3));
  const int stepShift = 2 + (j.imv == 3 ? 1 : (j.imv << 1));
  const int predX = change_prec(j.curPredX, down), predY = change_prec(j.curPredY, down);
  // removeWeightHighFreq's fixed-point weights (Buffer.cpp:236-262) and the distortion weight (:7666-7676)
  const int    bcwW   = j.bcwIdx == 0 ? -2 : (j.bcwIdx == 1 ? 3 : (j.bcwIdx == 2 ? 4 : (j.bcwIdx == 3 ? 5 : 10)));
  const int    norm   = ((1 << 16) + (bcwW > 0 ? (bcwW >> 1) : -(bcwW >> 1))) / bcwW;
  const int    weight0 = norm * 8, weight1 = (8 - bcwW) * norm;
  const double fWeight = j.bcwIdx == 2 ? 0.5 : fabs((double) bcwW / 8.0);
  // SATD tiling of the CU
  const int satdKind = !j.useHad ? 0 : (j.w > j.h ? 2 : (j.w < j.h ? 3 : 1));   // 1: 8x8, 2: 16x8, 3: 8x16

  if (threadIdx.x == 0)
  {
    st.curX = j.curMvX;
    st.curY = j.curMvY;
    st.tarX = j.tarMvX;
    st.tarY = j.tarMvY;
    st.cost = j.cost;
  }
  for (int phase = 0; phase < 2; phase++)
  {
    const int pattern = phase == 0 ? 2 : 0, maxRounds = phase == 0 ? (8 >> j.imv) : 1;
    const int rounding = pattern == 0 ? 4 : 8, mask = pattern == 0 ? 3 : 7;
    if (threadIdx.x == 0)
    {
      st.start = 0;
      st.end   = pattern == 0 ? 3 : 7;
      st.stop  = 0;
    }
    for (int round = 0; round < maxRounds; round++)
    {
      __syncthreads();
      if (st.stop) break;
      const int cx = st.curX, cy = st.curY, start = st.start, n = st.end - st.start + 1;
      if (threadIdx.x < 8) s_dist[threadIdx.x] = 0;
      __syncthreads();
      for (int u = warp; u < n * nTiles; u += kSmvdWarps)
      {
        const int c = u / nTiles, t = u - c * nTiles;
        const int direct = (start + c + rounding) & mask;
        int       ox, oy;
        if (pattern == 0)
        {
          ox = direct == 1 ? 1 : (direct == 3 ? -1 : 0);
          oy = direct == 0 ? 1 : (direct == 2 ? -1 : 0);
        }
        else
        {
          // { 0, 2 }, { 1, 1 }, { 2, 0 }, { 1, -1 }, { 0, -2 }, { -1, -1 }, { -2, 0 }, { -1, 1 }   (InterSearch.cpp:4398-4402)
          const int dxs[8] = { 0, 1, 2, 1, 0, -1, -2, -1 }, dys[8] = { 2, 1, 0, -1, -2, -1, 0, 1 };
          ox = dxs[direct];
          oy = dys[direct];
        }
        int mvAx = cx + ox * (1 << stepShift), mvAy = cy + oy * (1 << stepShift);
        int mvBx = j.tarPredX - (mvAx - j.curPredX), mvBy = j.tarPredY - (mvAy - j.curPredY);   // the mirrored MVD
        clip_mv(mvAx, mvAy, j);
        clip_mv(mvBx, mvBy, j);
        const int tx = (t % tilesX) * kTile, ty = (t / tilesX) * kTile;
        McTile    mt;
        mt.srcStride = j.refStride;
        mt.dst       = nullptr;
        mt.dstStride = 0;
        mt.tw        = (uint8_t) min(kTile, j.w - tx);
        mt.th        = (uint8_t) min(kTile, j.h - ty);
        mt.q4Hor = mt.q4Ver = 0;   // PUs of a symmetric-MVD search are at least 8 samples wide
        mt.winMaxX = mt.winMaxY = 0;
        mt.winX = mt.winY = 0;
        mt.src   = j.refCur + (ptrdiff_t) (j.y + ty + (mvAy >> 4)) * j.refStride + j.x + tx + (mvAx >> 4);
        mt.xFrac = (uint8_t) (mvAx & 15);
        mt.yFrac = (uint8_t) (mvAy & 15);
        TileSink sa{ s_predA[warp] };
        mc_tile<8>(mt, 0, j.bd, alt, s_patch[warp], s_mid[warp], lane, sa);
        __syncwarp();
        mt.src   = j.refTar + (ptrdiff_t) (j.y + ty + (mvBy >> 4)) * j.refStride + j.x + tx + (mvBx >> 4);
        mt.xFrac = (uint8_t) (mvBx & 15);
        mt.yFrac = (uint8_t) (mvBy & 15);
        TileSink sb{ s_predB[warp] };
        mc_tile<8>(mt, 0, j.bd, alt, s_patch[warp], s_mid[warp], lane, sb);
        __syncwarp();
        // difference (2*org - predA) - predB of this lane's samples
        const int16_t* org = j.org + (ptrdiff_t) ty * j.orgStride + tx;
        auto diff = [&](int y, int x) -> int
        {
          const int o = org[(ptrdiff_t) y * j.orgStride + x], a = s_predA[warp][y * kTile + x];
          int       v = j.bcwIdx == 2 ? 2 * o - a : (o * weight0 - a * weight1 + (1 << 15)) >> 16;
          if (j.clipBiPred) v = min(max(v, 0), maxv);
          return (int) (int16_t) v - (int) s_predB[warp][y * kTile + x];
        };
        const int tw = mt.tw, th = mt.th;
        uint32_t  part = 0;
        if (satdKind == 1)
        {
          const int  i = lane >> 3, r = lane & 7, bx = (i & 1) * 8, by = (i >> 1) * 8;
          const bool ok = bx < tw && by < th;
          int        d[8];
#pragma unroll
          for (int k = 0; k < 8; k++) d[k] = ok ? diff(by + r, bx + k) : 0;
          const uint32_t v = satd_tile_rows<8, 8>(d, r);
          part = (ok && r == 0) ? v : 0;
        }
        else if (satdKind == 2)   // 16 wide, 8 high: two tiles, lanes 0-7 and 8-15
        {
          const int  i = (lane >> 3) & 1, r = lane & 7, by = i * 8;
          const bool ok = lane < 16 && by < th;
          int        d[16];
#pragma unroll
          for (int k = 0; k < 16; k++) d[k] = ok ? diff(by + r, k) : 0;
          const uint32_t v = satd_tile_rows<16, 8>(d, r);
          part = (ok && r == 0) ? v : 0;
        }
        else if (satdKind == 3)   // 8 wide, 16 high: two tiles of 16 lanes
        {
          const int  i = lane >> 4, r = lane & 15, bx = i * 8;
          const bool ok = bx < tw;
          int        d[8];
#pragma unroll
          for (int k = 0; k < 8; k++) d[k] = ok ? diff(r, bx + k) : 0;
          const uint32_t v = satd_tile_rows<8, 16>(d, r);
          part = (ok && r == 0) ? v : 0;
        }
        else
        {
          for (int o = lane; o < tw * th; o += 32)
          {
            const int y = o / tw, x = o - y * tw;
            part += (uint32_t) abs(diff(y, x));
          }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
        if (lane == 0) atomicAdd(&s_dist[c], (unsigned long long) part);
        __syncwarp();
      }
      __syncthreads();
      if (threadIdx.x == 0)
      {
        // xSymmeticRefineMvSearch's loop over the candidates of this round (:4412-4470)
        int bestDirect = -1;
        for (int c = 0; c < n; c++)
        {
          const int direct = (start + c + rounding) & mask;
          int       ox, oy;
          if (pattern == 0)
          {
            ox = direct == 1 ? 1 : (direct == 3 ? -1 : 0);
            oy = direct == 0 ? 1 : (direct == 2 ? -1 : 0);
          }
          else
          {
            const int dxs[8] = { 0, 1, 2, 1, 0, -1, -2, -1 }, dys[8] = { 2, 1, 0, -1, -2, -1, 0, 1 };
            ox = dxs[direct];
            oy = dys[direct];
          }
          const int mx = cx + ox * (1 << stepShift), my = cy + oy * (1 << stepShift);
          const int px = j.tarPredX - (mx - j.curPredX), py = j.tarPredY - (my - j.curPredY);
          unsigned long long cost = mv_cost(j.lambda, mv_bits_q(change_prec(mx, down), change_prec(my, down), predX, predY, 0));
          cost += (unsigned long long) __dmul_rn(fWeight, (double) s_dist[c]);   // floor(fWeight * dist), dist >= 0
          if (cost < st.cost)
          {
            st.cost    = cost;
            st.curX    = mx;
            st.curY    = my;
            st.tarX    = px;
            st.tarY    = py;
            bestDirect = direct;
          }
        }
        if (bestDirect == -1) st.stop = 1;
        else
        {
          const int step = pattern == 2 ? 2 - (bestDirect & 1) : 1;
          st.start = bestDirect - step;
          st.end   = bestDirect + step;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    DevSmvdResult r;
    r.curMvX = st.curX;
    r.curMvY = st.curY;
    r.tarMvX = st.tarX;
    r.tarMvY = st.tarY;
    r.cost   = st.cost;
    out[blockIdx.x] = r;
  }
}

}   // namespace

cudaError_t launch_smvd_search(const DevSmvd* dJobs, DevSmvdResult* dOut, int n, cudaStream_t st)
{
  smvd_search_kernel<<<n, kSmvdWarps * 32, 0, st>>>(dJobs, dOut);
  return cudaGetLastError();
}

}   // namespace vtmme
