// Motion estimation of the GOP-based temporal filter, EncTemporalFilter::motionEstimation
// (EncoderLib/EncTemporalFilter.cpp:448-466): a four-level hierarchical block search of one reference frame against
// the original — 16x16 blocks on the 1/4 and 1/2 resolution pictures and on the full picture, then 8x8 blocks with a
// 1/16-sample refinement — with squared-error distortion (motionErrorLuma, :268-361) and a 6-tap interpolation.
//
//   mctf_subsample_kernel   subsampleLuma (:241-266): 2x2 mean, written with its replicated border in one pass
//   mctf_level_kernel       motionEstimationLuma (:363-446) for one level: one warp per block.  The reference tests its
//                           candidates one after the other with a strict "<"; between two decisions the candidate
//                           list is fixed (the 25 vectors inherited from the coarser level, the integer window around
//                           the best of those, the two 7x7 sub-sample grids), so each list is evaluated in parallel and
//                           reduced with the key (error, order): the first strict minimum in list order.  (The early
//                           exit of motionErrorLuma only ever returns values above the current best.)
//                           Inherited vectors: the warp walks them, lanes over the samples.  Window / sub-sample grids:
//                           one candidate per lane, the reference region staged in shared memory.
#include "me_kernels.h"

namespace vtmme {

namespace {

__constant__ int8_t c_mctfFilter[16][8] = {   // EncTemporalFilter::m_interpolationFilter (:50-68)
  { 0, 0, 0, 64, 0, 0, 0, 0 },    { 0, 1, -3, 64, 4, -2, 0, 0 },   { 0, 1, -6, 62, 9, -3, 1, 0 },   { 0, 2, -8, 60, 14, -5, 1, 0 },
  { 0, 2, -9, 57, 19, -7, 2, 0 }, { 0, 3, -10, 53, 24, -8, 2, 0 }, { 0, 3, -11, 50, 29, -9, 2, 0 }, { 0, 3, -11, 44, 35, -10, 3, 0 },
  { 0, 1, -7, 38, 38, -7, 1, 0 }, { 0, 3, -10, 35, 44, -11, 3, 0 }, { 0, 2, -9, 29, 50, -11, 3, 0 }, { 0, 2, -8, 24, 53, -10, 3, 0 },
  { 0, 2, -7, 19, 57, -9, 2, 0 }, { 0, 1, -5, 14, 60, -8, 2, 0 },  { 0, 1, -3, 9, 62, -6, 1, 0 },   { 0, 0, -2, 4, 64, -3, 1, 0 }
};

__global__ void mctf_subsample_kernel(DevPic in, DevPic out)
{
  const int W = out.width + 2 * out.margin, H = out.height + 2 * out.margin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < W * H; i += gridDim.x * blockDim.x)
  {
    const int y = i / W - out.margin, x = i % W - out.margin;
    const int sx = min(max(x, 0), out.width - 1), sy = min(max(y, 0), out.height - 1);
    const int16_t* a = in.origin + (ptrdiff_t) (2 * sy) * in.stride + 2 * sx;
    out.origin[(ptrdiff_t) y * out.stride + x] = (int16_t) ((a[0] + a[in.stride] + a[1] + a[in.stride + 1] + 2) >> 2);
  }
}

__global__ void mctf_init_mv_kernel(int3* mv, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mv[i] = make_int3(0, 0, 0x7fffffff);   // MotionVector() (EncTemporalFilter.h:50-56)
}

// EncTemporalFilter::applyMotion (:470-552) for one component: one thread per output sample.  The two passes of the
// reference keep exact integer sums between them, so a thread evaluates its own 6 x 6 support directly.
__global__ void mctf_apply_motion_kernel(DevPic src, int csx, int csy, const int3* __restrict__ mv, int mvStride, int maxv,
                                         int16_t* __restrict__ dst)
{
  const int bsx = 8 >> csx, bsy = 8 >> csy;
  const int wBlk = src.width / bsx * bsx, hBlk = src.height / bsy * bsy;   // whole blocks only (:492-494)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < src.width * src.height; i += gridDim.x * blockDim.x)
  {
    const int y = i / src.width, x = i - y * src.width;
    if (x >= wBlk || y >= hBlk)
    {
      dst[i] = 0;
      continue;
    }
    const int3 m  = mv[(size_t) (y / bsy) * mvStride + x / bsx];
    const int  dx = m.x >> csx, dy = m.y >> csy;
    const int8_t* xf = c_mctfFilter[dx & 15];
    const int8_t* yf = c_mctfFilter[dy & 15];
    const int16_t* p = src.origin + (ptrdiff_t) (y + (m.y >> (4 + csy)) - 2) * src.stride + (x + (m.x >> (4 + csx)) - 2);
    int sum = 0;
#pragma unroll
    for (int ky = 0; ky < 6; ky++)
    {
      int h = 0;
#pragma unroll
      for (int kx = 0; kx < 6; kx++) h += xf[kx + 1] * p[(ptrdiff_t) ky * src.stride + kx];
      sum += yf[ky + 1] * h;
    }
    sum    = (sum + (1 << 11)) >> 12;
    dst[i] = (int16_t) min(max(sum, 0), maxv);
  }
}

constexpr int kMctfWarps = 4;

template <int BS>
struct MctfSmem
{
  int16_t org[BS * BS];
  int16_t win[(BS + 16 + 2) * (BS + 16)];   // integer window: (BS + 2*range) rows of (BS + 2*range + 2) samples, range <= 8
};

// key of a candidate: (error, order in the list); errors stay below 2^29 (16x16 block, 10-bit samples)
__device__ __forceinline__ unsigned long long mctf_key(int error, int order)
{
  return ((unsigned long long) (uint32_t) error << 32) | (uint32_t) order;
}

// squared error of the BS x BS block against the reference at an integer displacement, read from global memory;
// the warp's lanes share the samples.  Returned in every lane.
template <int BS>
__device__ __forceinline__ int mctf_sse_global(const int16_t* s_org, const int16_t* ref, int refStride)
{
  const int lane = threadIdx.x & 31;
  int       e    = 0;
#pragma unroll
  for (int i = lane; i < BS * BS; i += 32)
  {
    const int r = i / BS, c = i % BS;
    const int d = (int) s_org[i] - (int) ref[(ptrdiff_t) r * refStride + c];
    e += d * d;
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) e += __shfl_xor_sync(0xffffffffu, e, m);
  return e;
}

// squared error at displacement (dx, dy) in 1/16 sample, one lane, reference read from global memory: the 6-tap
// separable interpolation of motionErrorLuma (:309-357).  The horizontal sums are exact integers, the vertical pass
// rounds once ((sum + 2048) >> 12) and clips; phase 0 reproduces the integer path exactly.
template <int BS>
__device__ __forceinline__ int mctf_sse_frac(const int16_t* s_org, const int16_t* refAtBlock, int refStride, int dx, int dy,
                                             int maxv)
{
  int xf[6], yf[6];
#pragma unroll
  for (int k = 0; k < 6; k++)
  {
    xf[k] = c_mctfFilter[dx & 15][k + 1];
    yf[k] = c_mctfFilter[dy & 15][k + 1];
  }
  const int16_t* base = refAtBlock + (ptrdiff_t) ((dy >> 4) - 2) * refStride + ((dx >> 4) - 2);   // tap k = 1 of row y1 = 1
  int            ring[6][BS];   // horizontally filtered rows y1 .. y1 + 5 (statically indexed: the loops unroll)
  int            e = 0;
#pragma unroll
  for (int y1 = 0; y1 < BS + 5; y1++)
  {
    const int16_t* row = base + (ptrdiff_t) y1 * refStride;
    int            p[BS + 5];
#pragma unroll
    for (int i = 0; i < BS + 5; i++) p[i] = row[i];
    int h[BS];
#pragma unroll
    for (int x1 = 0; x1 < BS; x1++)
    {
      int s = 0;
#pragma unroll
      for (int k = 0; k < 6; k++) s += xf[k] * p[x1 + k];
      h[x1] = s;
    }
#pragma unroll
    for (int x1 = 0; x1 < BS; x1++) ring[y1 % 6][x1] = h[x1];
    if (y1 >= 5)
    {
      const int oy = y1 - 5;   // output row: taps rows oy .. oy + 5
#pragma unroll
      for (int x1 = 0; x1 < BS; x1++)
      {
        int s = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) s += yf[k] * ring[(oy + k) % 6][x1];
        s = (s + (1 << 11)) >> 12;
        s = min(max(s, 0), maxv);
        const int d = s - (int) s_org[oy * BS + x1];
        e += d * d;
      }
    }
  }
  return e;
}

template <int BS, bool DOUBLE>
__global__ void __launch_bounds__(kMctfWarps * 32) mctf_level_kernel(MctfLevelParams p)
{
  __shared__ MctfSmem<BS> s_all[kMctfWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbx = (p.width - 1) / BS, nby = (p.height - 1) / BS;   // blocks with blockX + BS < width (:372-374)
  const int blk = blockIdx.x * kMctfWarps + warp, pair = blockIdx.y;
  if (blk >= nbx * nby) return;
  MctfSmem<BS>& sm = s_all[warp];
  const int bx = blk % nbx, by = blk / nbx, blockX = bx * BS, blockY = by * BS;
  const DevPic org = p.org[pair], ref = p.ref[pair];
  const int16_t* orgBlk = org.origin + (ptrdiff_t) blockY * org.stride + blockX;
  const int16_t* refBlk = ref.origin + (ptrdiff_t) blockY * ref.stride + blockX;
#pragma unroll
  for (int i = lane; i < BS * BS; i += 32) sm.org[i] = orgBlk[(ptrdiff_t) (i / BS) * org.stride + (i % BS)];
  __syncwarp();

  int bestX = 0, bestY = 0, bestE = 0x7fffffff;
  int range = 8;
  if (p.previous)
  {
    // the 25 vectors of the coarser level around the block (:383-401), in raster order
    range = 5;
    const int3* prev = p.previous + (size_t) pair * p.prevW * p.prevH;
    const int   cx = blockX / (2 * BS), cy = blockY / (2 * BS);
    const int   limX = p.width / (2 * BS), limY = p.height / (2 * BS);
    for (int py = -2; py <= 2; py++)
      for (int px = -2; px <= 2; px++)
      {
        const int tx = cx + px, ty = cy + py;
        if (tx < 0 || tx >= limX || ty < 0 || ty >= limY) continue;   // uniform across the warp
        const int3 old = prev[ty * p.prevW + tx];
        const int  mx = old.x * p.factor, my = old.y * p.factor;
        // inherited vectors are whole samples: only the last level refines below one sample
        const int e = mctf_sse_global<BS>(sm.org, refBlk + (ptrdiff_t) (my / 16) * ref.stride + mx / 16, ref.stride);
        if (e < bestE)
        {
          bestE = e;
          bestX = mx;
          bestY = my;
        }
      }
  }
  {
    // integer window around the best so far (:402-414): (2 range + 1)^2 candidates, one per lane, region in shared memory
    const int cxi = bestX / 16, cyi = bestY / 16;   // C division, like the reference
    const int span = 2 * range + 1, ww = BS + 2 * range, pitch = BS + 16 + 2;
    for (int i = lane; i < ww * ww; i += 32)
    {
      const int r = i / ww, c = i % ww;
      sm.win[r * pitch + c] = refBlk[(ptrdiff_t) (cyi - range + r) * ref.stride + (cxi - range + c)];
    }
    __syncwarp();
    unsigned long long k = ~0ull;
    for (int c0 = 0; c0 < span * span; c0 += 32)
    {
      const int c = c0 + lane;
      if (c < span * span)
      {
        const int      wy = c / span, wx = c % span;
        const int16_t* w  = sm.win + wy * pitch + wx;
        int            e  = 0;
#pragma unroll 4
        for (int r = 0; r < BS; r++)
#pragma unroll
          for (int x = 0; x < BS; x++)
          {
            const int d = (int) sm.org[r * BS + x] - (int) w[r * pitch + x];
            e += d * d;
          }
        const unsigned long long kc = mctf_key(e, c);
        k = kc < k ? kc : k;
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1)
    {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, m);
      k                          = o < k ? o : k;
    }
    const int e = (int) (k >> 32), c = (int) (uint32_t) k;
    if (e < bestE)
    {
      bestE = e;
      bestX = (cxi - range + c % span) * 16;
      bestY = (cyi - range + c / span) * 16;
    }
    __syncwarp();
  }
  if (DOUBLE)
  {
    // +-12/16 in steps of 4/16, then +-3/16 in steps of 1/16 (:415-441): 49 candidates per pass, one per lane
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++)
    {
      const int dr = pass ? 3 : 12, st = pass ? 1 : 4;
      const int px = bestX, py = bestY;
      unsigned long long k = ~0ull;
#pragma unroll 1
      for (int c0 = 0; c0 < 49; c0 += 32)
      {
        const int c = c0 + lane;
        if (c < 49)
        {
          const int dx = px - dr + (c % 7) * st, dy = py - dr + (c / 7) * st;
          const int e  = mctf_sse_frac<BS>(sm.org, refBlk, ref.stride, dx, dy, p.maxv);
          const unsigned long long kc = mctf_key(e, c);
          k = kc < k ? kc : k;
        }
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1)
      {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, m);
        k                          = o < k ? o : k;
      }
      const int e = (int) (k >> 32), c = (int) (uint32_t) k;
      if (e < bestE)
      {
        bestE = e;
        bestX = px - dr + (c % 7) * st;
        bestY = py - dr + (c / 7) * st;
      }
    }
  }
  if (lane == 0) p.mvs[(size_t) pair * p.mvW * p.mvH + (size_t) by * p.mvW + bx] = make_int3(bestX, bestY, bestE);
}

// EncTemporalFilter::bilateralFilter, the weighting of one component (EncoderLib/EncTemporalFilter.cpp:591-618).  The weight
// of a neighbouring sample, weightScaling * refStrength * exp(-(diff * 1024 / 2^bd)^2 / (2 sigma^2)), depends only on
// |refVal - orgVal| and the picture's POC-distance class, so the host hands over one table of 2^bd doubles per neighbouring
// picture (computed with its own exp(), the reference's libm); products and sums are IEEE doubles in the reference's order,
// never contracted into FMAs, one division, round() half away from zero: every output equals the reference's.
struct MctfBilateralParams
{
  DevPic        org;
  DevPic        corr[8];
  const double* weights;   // [numRefs][1 << bitDepth]
  int           numRefs, tableSize, maxv;
  int16_t*      dst;       // packed width x height
};

__global__ void __launch_bounds__(256) mctf_bilateral_kernel(MctfBilateralParams p)
{
  const int       w = p.org.width, h = p.org.height;
  const long long n = (long long) w * h, stride = (long long) gridDim.x * blockDim.x;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const int y = (int) (i / w), x = (int) (i - (long long) y * w);
    const int orgVal = p.org.origin[(size_t) y * p.org.stride + x];
    double    sum = 1.0, newVal = (double) orgVal;
    for (int r = 0; r < p.numRefs; r++)
    {
      const int    refVal = p.corr[r].origin[(size_t) y * p.corr[r].stride + x];
      const int    d      = min(abs(refVal - orgVal), p.tableSize - 1);
      const double wgt    = p.weights[(size_t) r * p.tableSize + d];
      newVal = __dadd_rn(newVal, __dmul_rn(wgt, (double) refVal));
      sum    = __dadd_rn(sum, wgt);
    }
    const int v = (int) (short) (int) round(__ddiv_rn(newVal, sum));
    p.dst[i]    = (int16_t) min(max(v, 0), p.maxv);
  }
}

}   // namespace

cudaError_t launch_mctf_bilateral(DevPic org, const DevPic* corr, int numRefs, const double* dWeights, int bitDepth, int16_t* dst,
                                  cudaStream_t st)
{
  MctfBilateralParams p;
  p.org = org;
  for (int i = 0; i < numRefs; i++) p.corr[i] = corr[i];
  p.weights   = dWeights;
  p.numRefs   = numRefs;
  p.tableSize = 1 << bitDepth;
  p.maxv      = (1 << bitDepth) - 1;
  p.dst       = dst;
  mctf_bilateral_kernel<<<592, 256, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_mctf_subsample(DevPic in, DevPic out, cudaStream_t st)
{
  mctf_subsample_kernel<<<296, 256, 0, st>>>(in, out);
  return cudaGetLastError();
}

cudaError_t launch_mctf_apply_motion(DevPic src, int csx, int csy, const int3* mv, int mvStride, int maxv, int16_t* dst,
                                     cudaStream_t st)
{
  mctf_apply_motion_kernel<<<592, 256, 0, st>>>(src, csx, csy, mv, mvStride, maxv, dst);
  return cudaGetLastError();
}

cudaError_t launch_mctf_init_mv(int3* mv, int n, cudaStream_t st)
{
  mctf_init_mv_kernel<<<(n + 255) / 256, 256, 0, st>>>(mv, n);
  return cudaGetLastError();
}

cudaError_t launch_mctf_level(const MctfLevelParams& p, int blockSize, bool doubleRes, int nPairs, cudaStream_t st)
{
  const int nb = ((p.width - 1) / blockSize) * ((p.height - 1) / blockSize);
  if (nb <= 0) return cudaSuccess;
  dim3 grid((nb + kMctfWarps - 1) / kMctfWarps, nPairs, 1);
  if (blockSize == 16 && !doubleRes)
    mctf_level_kernel<16, false><<<grid, kMctfWarps * 32, 0, st>>>(p);
  else if (blockSize == 8 && doubleRes)
    mctf_level_kernel<8, true><<<grid, kMctfWarps * 32, 0, st>>>(p);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}   // namespace vtmme
