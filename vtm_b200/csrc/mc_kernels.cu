// Motion compensation of a batch of blocks and the two element-wise helpers of bi-prediction.
//
//   InterPrediction::xPredInterBlk   CommonLib/InterPrediction.cpp:660-830 (plain path: no RPR, wrap-around, BDOF padding,
//                                    DMVR or bilinear filters)
//   AreaBuf<Pel>::addAvg             CommonLib/Buffer.cpp:467-507
//   AreaBuf<T>::removeHighFreq       CommonLib/Buffer.h:474-517
//
// HBM/L2-bound work: a block is cut into tiles of at most 16x16 outputs, one warp per tile.  The warp stages the tile's
// reference patch (with the tap halo the fractional phases need) in shared memory with row-major loads, runs the
// horizontal pass into a 14-bit intermediate plane and the vertical pass out of it, and writes the prediction rows.
// Warps are independent (__syncwarp only), so ragged batches of mixed block sizes keep every warp busy.
#include "mc_tile.cuh"
#include "me_kernels.h"

namespace vtmme {
namespace {

constexpr int kMcWarps    = 8;
using namespace mc;

struct StoreSink
{
  int16_t* dst;
  int      stride;
  __device__ __forceinline__ void operator()(int y, int x, int v) const { dst[(size_t) y * stride + x] = (int16_t) v; }
};

template <int TAPS>
__global__ void __launch_bounds__(kMcWarps * 32) mc_batch_kernel(const McTile* __restrict__ tiles, int nTiles, int bi,
                                                                  int bitDepth, int alt)
{
  __shared__ int16_t s_patch[kMcWarps][kPatchRows * kPatchPitch];
  __shared__ int16_t s_mid[kMcWarps][kPatchRows * kTile];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx  = blockIdx.x * kMcWarps + warp;
  if (idx >= nTiles) return;   // warps are independent: no CTA-wide barrier below
  const McTile t = tiles[idx];
  StoreSink    sink{ t.dst, t.dstStride };
  mc_tile<TAPS>(t, bi, bitDepth, alt, s_patch[warp], s_mid[warp], lane, sink);
}

// SAD of the original block against the motion-compensated prediction at a (fractional) candidate MV, without
// materialising the prediction: RdCost::xGetSAD (RdCost.cpp:493-528) on xPredInterBlk's output — the distortion term of
// InterSearch::xGetTemplateCost (InterSearch.cpp:3235-3270) and, for integer MVs, of the ME seeds (:3388-3426).
struct SadSink
{
  const int16_t* org;
  int            stride, rowMask;
  uint32_t       acc;
  __device__ __forceinline__ void operator()(int y, int x, int v)
  {
    if ((y & rowMask) == 0) acc += (uint32_t) abs((int) org[(size_t) y * stride + x] - v);
  }
};

__global__ void __launch_bounds__(kMcWarps * 32) mc_sad_kernel(const McSadTile* __restrict__ tiles, int nTiles, int bitDepth,
                                                                int alt, unsigned long long* __restrict__ out)
{
  __shared__ int16_t s_patch[kMcWarps][kPatchRows * kPatchPitch];
  __shared__ int16_t s_mid[kMcWarps][kPatchRows * kTile];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx  = blockIdx.x * kMcWarps + warp;
  if (idx >= nTiles) return;
  const McSadTile t = tiles[idx];
  SadSink         sink{ t.org, t.orgStride, (1 << t.subShift) - 1, 0u };
  mc_tile<8>(t.mc, 0, bitDepth, alt, s_patch[warp], s_mid[warp], lane, sink);
  uint32_t s = sink.acc;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (lane == 0) atomicAdd(out + t.outIdx, (unsigned long long) s << t.subShift);
}

// addAvg: (a + b + offset) >> shift, clipped — the default bi-prediction average of two 14-bit predictions
__global__ void __launch_bounds__(256) add_avg_kernel(const int16_t* __restrict__ a, const int16_t* __restrict__ b,
                                                      int16_t* __restrict__ d, long long n, int shift, int offset, int maxv)
{
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const int v = ((int) a[i] + (int) b[i] + offset) >> shift;
    d[i]        = (int16_t) min(max(v, 0), maxv);
  }
}

// the same on 8 samples per thread (16-byte accesses) for aligned buffers
__global__ void __launch_bounds__(256) add_avg_vec_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                          uint4* __restrict__ d, long long nVec, int shift, int offset, int maxv)
{
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < nVec; i += stride)
  {
    const uint4 va = a[i], vb = b[i];
    const uint32_t wa[4] = { va.x, va.y, va.z, va.w }, wb[4] = { vb.x, vb.y, vb.z, vb.w };
    uint32_t       wd[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
      const int lo = ((int) (int16_t) (wa[k] & 0xffffu) + (int) (int16_t) (wb[k] & 0xffffu) + offset) >> shift;
      const int hi = ((int) (int16_t) (wa[k] >> 16) + (int) (int16_t) (wb[k] >> 16) + offset) >> shift;
      wd[k] = (uint32_t) min(max(lo, 0), maxv) | ((uint32_t) min(max(hi, 0), maxv) << 16);
    }
    d[i] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
  }
}

// removeHighFreq: dst = 2*dst - src (optionally clipped): the search target of the second bi-prediction direction
__global__ void __launch_bounds__(256) remove_high_freq_kernel(int16_t* __restrict__ d, const int16_t* __restrict__ s,
                                                               long long n, int clip, int maxv)
{
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    int v = 2 * (int) d[i] - (int) s[i];
    if (clip) v = min(max(v, 0), maxv);
    d[i] = (int16_t) v;
  }
}

// BCW forms (Buffer.cpp addWeightedAvg / Buffer.h:124 removeWeightHighFreq).  addWeightedAvg: (a*w0 + b*w1 + offset) >> shift,
// clipped; removeWeightHighFreq: (org*weight0 - pred*weight1 + 2^15) >> 16 with the 16-bit normaliser of the BCW weight.
__global__ void __launch_bounds__(256) add_weighted_avg_kernel(const int16_t* __restrict__ a, const int16_t* __restrict__ b,
                                                               int16_t* __restrict__ d, long long n, int w0, int w1, int shift,
                                                               int offset, int maxv)
{
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const int v = ((int) a[i] * w0 + (int) b[i] * w1 + offset) >> shift;
    d[i]        = (int16_t) min(max(v, 0), maxv);
  }
}

__global__ void __launch_bounds__(256) remove_weight_high_freq_kernel(int16_t* __restrict__ d, const int16_t* __restrict__ s,
                                                                      long long n, int clip, int maxv, int weight0, int weight1)
{
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    int v = ((int) d[i] * weight0 - (int) s[i] * weight1 + (1 << 15)) >> 16;
    if (clip) v = min(max(v, 0), maxv);
    d[i] = (int16_t) v;
  }
}

int elementwise_grid(long long n)
{
  long long g = (n + 255) / 256;
  return (int) (g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}   // namespace

cudaError_t launch_mc_batch(int comp, const McTile* dTiles, int nTiles, int bi, int bitDepth, int useAltHpel, cudaStream_t st)
{
  const int grid = (nTiles + kMcWarps - 1) / kMcWarps;
  if (comp == 0)
    mc_batch_kernel<8><<<grid, kMcWarps * 32, 0, st>>>(dTiles, nTiles, bi, bitDepth, useAltHpel);
  else
    mc_batch_kernel<4><<<grid, kMcWarps * 32, 0, st>>>(dTiles, nTiles, bi, bitDepth, 0);
  return cudaGetLastError();
}

cudaError_t launch_mc_sad(const McSadTile* dTiles, int nTiles, int bitDepth, int useAltHpel, unsigned long long* dOut,
                          cudaStream_t st)
{
  mc_sad_kernel<<<(nTiles + kMcWarps - 1) / kMcWarps, kMcWarps * 32, 0, st>>>(dTiles, nTiles, bitDepth, useAltHpel, dOut);
  return cudaGetLastError();
}

cudaError_t launch_add_avg(const int16_t* a, const int16_t* b, int16_t* d, long long n, int bitDepth, cudaStream_t st)
{
  const int hr = 14 - bitDepth > 2 ? 14 - bitDepth : 2, shift = hr + 1, offset = (1 << (shift - 1)) + 2 * 8192;
  const int maxv = (1 << bitDepth) - 1;
  const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
  const long long nVec = aligned ? n / 8 : 0;
  if (nVec)
    add_avg_vec_kernel<<<elementwise_grid(nVec), 256, 0, st>>>(reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b),
                                                               reinterpret_cast<uint4*>(d), nVec, shift, offset, maxv);
  if (n - nVec * 8)
    add_avg_kernel<<<elementwise_grid(n - nVec * 8), 256, 0, st>>>(a + nVec * 8, b + nVec * 8, d + nVec * 8, n - nVec * 8, shift,
                                                                    offset, maxv);
  return cudaGetLastError();
}

cudaError_t launch_add_weighted_avg(const int16_t* a, const int16_t* b, int16_t* d, long long n, int bitDepth, int bcwIdx,
                                    cudaStream_t st)
{
  static const int bcwWeights[5] = { -2, 3, 4, 5, 10 };   // g_BcwWeights (CommonLib/Rom.cpp), BCW_DEFAULT = 2
  const int w1 = bcwWeights[bcwIdx], w0 = 8 - w1;
  const int hr = 14 - bitDepth > 2 ? 14 - bitDepth : 2, shift = hr + 3, offset = (1 << (shift - 1)) + (8192 << 3);
  add_weighted_avg_kernel<<<elementwise_grid(n), 256, 0, st>>>(a, b, d, n, w0, w1, shift, offset, (1 << bitDepth) - 1);
  return cudaGetLastError();
}

cudaError_t launch_remove_weight_high_freq(int16_t* d, const int16_t* s, long long n, int clip, int bitDepth, int bcwWeight,
                                           cudaStream_t st)
{
  const int normalizer = ((1 << 16) + (bcwWeight > 0 ? (bcwWeight >> 1) : -(bcwWeight >> 1))) / bcwWeight;
  remove_weight_high_freq_kernel<<<elementwise_grid(n), 256, 0, st>>>(d, s, n, clip, (1 << bitDepth) - 1, normalizer * 8,
                                                                      (8 - bcwWeight) * normalizer);
  return cudaGetLastError();
}

cudaError_t launch_remove_high_freq(int16_t* d, const int16_t* s, long long n, int clip, int bitDepth, cudaStream_t st)
{
  remove_high_freq_kernel<<<elementwise_grid(n), 256, 0, st>>>(d, s, n, clip, (1 << bitDepth) - 1);
  return cudaGetLastError();
}

}   // namespace vtmme
