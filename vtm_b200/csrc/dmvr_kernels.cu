// Decoder-side MV refinement (DMVR): the search of InterPrediction::xProcessDMVR (CommonLib/InterPrediction.cpp:2098-2154)
// for a batch of sub-blocks — bilinear prediction of the (w+4) x (h+4) neighbourhood of both lists (xPrefetch :1664-1708,
// xinitMC :1949-1995, xPredInterBlk with bilinearMC on m_bilinearFilterPrec4 and the biMCForDMVR rounding,
// InterpolationFilter.cpp:312-330,424-452,600-612), SAD of the even rows at the 25 mirrored integer offsets (xDMVRCost
// :1919-1927, xBIPMVRefine :1820-1843) and the parametric sub-sample step (xDMVRSubPixelErrorSurface :1929-1947,
// xSubPelErrorSrfc :1766-1818, div_for_maxq7 :1731-1763).
//
// One warp per sub-block, four per CTA, warp-level synchronisation only.  The reference computes the 25 costs one after
// the other and keeps the first strict minimum in the order of m_pSearchOffset; the offsets are fixed, so the costs are
// computed in parallel (one offset per lane) and one lane replays the decision.  Sub-blocks are independent: a whole
// picture's bi-predicted merge blocks are one launch.
#include "me_common.cuh"
#include "me_kernels.h"

namespace vtmme {

namespace {

constexpr int kDmvrWarps = 4;
constexpr int kDmvrSide  = 20;   // 16 + 2 * DMVR_NUM_ITERATION

struct DmvrSmem
{
  int16_t  pred[2][kDmvrSide * kDmvrSide];        // bilinear predictions of list 0 / list 1, row stride w + 4
  int16_t  hor[(kDmvrSide + 1) * kDmvrSide];      // horizontally filtered rows of the two-pass case
  uint32_t sad[25];
};

// clipMvInPic (Mv.cpp:53-71)
__device__ __forceinline__ void dmvr_clip(int& mx, int& my, int x, int y, int picW, int picH, int maxCu)
{
  mx = clampi(mx, (-maxCu - 8 - x + 1) * 16, (picW + 8 - x - 1) * 16);
  my = clampi(my, (-maxCu - 8 - y + 1) * 16, (picH + 8 - y - 1) * 16);
}

__device__ __forceinline__ void dmvr_bilinear(const DevPic& ref, int x, int y, int w, int h, int mvx, int mvy, int maxCu, int bd,
                                              int16_t* dst, int16_t* hor, int lane)
{
  const int W2 = w + 4, H2 = h + 4;
  int       cx = mvx - 48, cy = mvy - 48, fx = mvx, fy = mvy;
  dmvr_clip(cx, cy, x, y, ref.width, ref.height, maxCu);   // xPrefetch: where the samples are fetched
  dmvr_clip(fx, fy, x, y, ref.width, ref.height, maxCu);   // xinitMC: the fraction
  const int16_t* src = ref.origin + (ptrdiff_t) (y + (cy >> 4) + 1) * ref.stride + x + (cx >> 4) + 1;
  const int      xf = fx & 15, yf = fy & 15;
  const int      sh1 = 4 - (10 - bd), off1 = 1 << (sh1 - 1);
  // horizontal stage over H2 + 1 rows (raw samples when xf == 0)
  for (int i = lane; i < (H2 + 1) * W2; i += 32)
  {
    const int      r = i / W2, c = i - r * W2;
    const int16_t* s = src + (ptrdiff_t) r * ref.stride + c;
    hor[i]           = xf ? (int16_t) ((s[0] * (16 - xf) + s[1] * xf + off1) >> sh1) : s[0];
  }
  __syncwarp();
  for (int i = lane; i < H2 * W2; i += 32)
  {
    int v;
    if (yf == 0)
      v = xf ? hor[i] : hor[i] << (10 - bd);
    else if (xf == 0)
      v = (hor[i] * (16 - yf) + hor[i + W2] * yf + off1) >> sh1;
    else
      v = (hor[i] * (16 - yf) + hor[i + W2] * yf + 8) >> 4;
    dst[i] = (int16_t) v;
  }
  __syncwarp();
}

__device__ __forceinline__ int dmvr_div_maxq7(long long num, long long den)
{
  bool neg = false;
  int  q   = 0;
  if (num < 0)
  {
    neg = true;
    num = -num;
  }
  den <<= 3;
  if (num >= den)
  {
    num -= den;
    q++;
  }
  q <<= 1;
  den >>= 1;
  if (num >= den)
  {
    num -= den;
    q++;
  }
  q <<= 1;
  if (num >= (den >> 1)) q++;
  return neg ? -q : q;
}

__global__ void __launch_bounds__(kDmvrWarps * 32) dmvr_refine_kernel(DevPic ref0, DevPic ref1, const DevDmvrBlock* __restrict__ blocks,
                                                                       int n, int bd, int maxCu, DevDmvrResult* __restrict__ results)
{
  __shared__ DmvrSmem sm[kDmvrWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bi = blockIdx.x * kDmvrWarps + warp;
  if (bi >= n) return;
  DmvrSmem&          s = sm[warp];
  const DevDmvrBlock b = blocks[bi];
  const int          W2 = b.w + 4;
  dmvr_bilinear(ref0, b.x, b.y, b.w, b.h, b.mv0x, b.mv0y, maxCu, bd, s.pred[0], s.hor, lane);
  dmvr_bilinear(ref1, b.x, b.y, b.w, b.h, b.mv1x, b.mv1y, maxCu, bd, s.pred[1], s.hor, lane);
  if (lane < 25)
  {
    // offset (ox, oy) on list 0, mirrored on list 1; even rows (subShift 1, its << 1 undone by xDMVRCost)
    const int ox = lane % 5 - 2, oy = lane / 5 - 2;
    uint32_t  acc = 0;
    for (int r = 0; r < b.h; r += 2)
    {
      const int16_t* p0 = s.pred[0] + (2 + oy + r) * W2 + 2 + ox;
      const int16_t* p1 = s.pred[1] + (2 - oy + r) * W2 + 2 - ox;
      for (int c = 0; c < b.w; c++) acc = __sad((int) p0[c], (int) p1[c], acc);
    }
    s.sad[lane] = acc;
  }
  __syncwarp();
  if (lane == 0)
  {
    uint32_t minCost = s.sad[12] - (s.sad[12] >> 2);
    int      bestK = 12, notZero = 1;
    if (minCost < (uint32_t) (b.w * b.h))
      notZero = 0;
    else
    {
      s.sad[12] = minCost;
      if (!minCost)
        notZero = 0;
      else
        for (int k = 0; k < 25; k++)
          if (s.sad[k] < minCost)
          {
            minCost = s.sad[k];
            bestK   = k;
          }
    }
    int tx = (bestK % 5 - 2) * 16, ty = (bestK / 5 - 2) * 16;
    if (notZero && abs(tx) != 32 && abs(ty) != 32)
    {
      const long long e0 = s.sad[bestK], e1 = s.sad[bestK - 1], e2 = s.sad[bestK - 5], e3 = s.sad[bestK + 1], e4 = s.sad[bestK + 5];
      long long       den = e1 + e3 - (e0 << 1);
      if (den != 0) tx += (e1 != e0 && e3 != e0) ? dmvr_div_maxq7((e1 - e3) << 4, den) : (e1 == e0 ? -8 : 8);
      den = e2 + e4 - (e0 << 1);
      if (den != 0) ty += (e2 != e0 && e4 != e0) ? dmvr_div_maxq7((e2 - e4) << 4, den) : (e2 == e0 ? -8 : 8);
    }
    DevDmvrResult r;
    r.mvdX        = tx;
    r.mvdY        = ty;
    r.minCost     = minCost;
    r.notZeroCost = notZero;
    results[bi]   = r;
  }
}

}   // namespace

cudaError_t launch_dmvr_refine(const DevPic& ref0, const DevPic& ref1, const DevDmvrBlock* blocks, int n, int bitDepth, int maxCu,
                               DevDmvrResult* results, cudaStream_t st)
{
  dmvr_refine_kernel<<<(n + kDmvrWarps - 1) / kDmvrWarps, kDmvrWarps * 32, 0, st>>>(ref0, ref1, blocks, n, bitDepth, maxCu, results);
  return cudaGetLastError();
}

}   // namespace vtmme
