// Affine motion estimation: the three dispatch-table primitives of AffineGradientSearch
// (CommonLib/AffineGradientSearch.cpp:64-174, SIMD in x86/AffineGradientSearchX86.h) that every iteration of
// InterSearch::xAffineMotionEstimation calls (EncoderLib/InterSearch.cpp:5486-5526): the 3x3 Sobel derivatives of the current
// prediction (border = nearest interior value) and the normal-equation sums of xEqualCoeffComputer — and their fusion into one
// launch per batch of blocks (error = org - pred, both derivatives on the fly, int64 sums), which is what a batched caller
// needs per iteration.  All integer arithmetic: derivatives fit 14 bits, the products 32 bits, the sums are 64-bit.
#include "me_kernels.h"

namespace vtmme {
namespace {

constexpr int kAffThreads = 256;

__device__ __forceinline__ int sobel_at(const int16_t* __restrict__ pred, int stride, int w, int h, int j, int k, bool vertical)
{
  const int      jj = min(max(j, 1), h - 2), kk = min(max(k, 1), w - 2);
  const int16_t* c  = pred + (ptrdiff_t) jj * stride + kk;
  if (!vertical) return (int) c[1 - stride] - (int) c[-1 - stride] + ((int) c[1] << 1) - ((int) c[-1] << 1) + (int) c[1 + stride] - (int) c[-1 + stride];
  return (int) c[stride - 1] - (int) c[-stride - 1] + ((int) c[stride] << 1) - ((int) c[-stride] << 1) + (int) c[stride + 1] - (int) c[-stride + 1];
}

__global__ void __launch_bounds__(kAffThreads) affine_sobel_kernel(const int16_t* __restrict__ pred, int stride, int w, int h, int vertical,
                                                                   int* __restrict__ deriv)
{
  for (int i = blockIdx.x * kAffThreads + threadIdx.x; i < w * h; i += gridDim.x * kAffThreads)
    deriv[i] = sobel_at(pred, stride, w, h, i / w, i % w, vertical != 0);
}

// the upper triangle of sum(c_col * c_row) and the residual column, per thread; SIX: 6-parameter model
template <bool SIX>
struct CoeffAcc
{
  static constexpr int N = SIX ? 6 : 4, NT = N * (N + 1) / 2 + N;
  long long v[NT];
  __device__ __forceinline__ void clear()
  {
#pragma unroll
    for (int i = 0; i < NT; i++) v[i] = 0;
  }
  __device__ __forceinline__ void add(int j, int k, int gx, int gy, int residue)
  {
    const int cx = ((k >> 2) << 2) + 2, cy = ((j >> 2) << 2) + 2;
    int       c[N];
    if (!SIX)
    {
      c[0] = gx;
      c[1] = cx * gx + cy * gy;
      c[2] = gy;
      c[3] = cy * gx - cx * gy;
    }
    else
    {
      c[0] = gx;
      c[1] = cx * gx;
      c[2] = gy;
      c[3] = cx * gy;
      c[N - 2] = cy * gx;
      c[N - 1] = cy * gy;
    }
    int t = 0;
#pragma unroll
    for (int col = 0; col < N; col++)
    {
#pragma unroll
      for (int row = col; row < N; row++) v[t++] += (long long) c[col] * c[row];
      v[t++] += ((long long) c[col] * residue) << 3;
    }
  }
  // block-wide sum into coeff[7][7] (accumulating, like the reference): rows 1..N, columns 0..N
  __device__ __forceinline__ void reduce_into(long long* coeff, unsigned long long* s_sum)
  {
    for (int i = threadIdx.x; i < NT; i += kAffThreads) s_sum[i] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NT; i++)
    {
      long long x = v[i];
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_sum[i], (unsigned long long) x);
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
      int t = 0;
      for (int col = 0; col < N; col++)
      {
        for (int row = col; row < N; row++)
        {
          const long long s = (long long) s_sum[t++];
          coeff[(col + 1) * 7 + row] += s;
          if (row != col) coeff[(row + 1) * 7 + col] += s;
        }
        coeff[(col + 1) * 7 + N] += (long long) s_sum[t++];
      }
    }
  }
};

template <bool SIX>
__global__ void __launch_bounds__(kAffThreads) affine_equal_coeff_kernel(const int16_t* __restrict__ residue, int residueStride,
                                                                         const int* __restrict__ d0, const int* __restrict__ d1, int derivStride,
                                                                         int w, int h, long long* coeff)
{
  __shared__ unsigned long long s_sum[32];
  CoeffAcc<SIX> acc;
  acc.clear();
  for (int i = threadIdx.x; i < w * h; i += kAffThreads)
  {
    const int j = i / w, k = i - j * w;
    acc.add(j, k, d0[(size_t) j * derivStride + k], d1[(size_t) j * derivStride + k], residue[(size_t) j * residueStride + k]);
  }
  acc.reduce_into(coeff, s_sum);
}

// one iteration's gradient step for a batch of blocks: error, Sobel derivatives and sums in one pass; one CTA per block
template <bool SIX>
__device__ __forceinline__ void affine_step_block(const DevAffineBlock& b, long long* coeff, unsigned long long* s_sum)
{
  CoeffAcc<SIX> acc;
  acc.clear();
  for (int i = threadIdx.x; i < b.w * b.h; i += kAffThreads)
  {
    const int j = i / b.w, k = i - j * b.w;
    const int gx = sobel_at(b.pred, b.predStride, b.w, b.h, j, k, false), gy = sobel_at(b.pred, b.predStride, b.w, b.h, j, k, true);
    const int e  = (int) (int16_t) ((int) b.org[(size_t) j * b.orgStride + k] - (int) b.pred[(size_t) j * b.predStride + k]);
    acc.add(j, k, gx, gy, e);
  }
  acc.reduce_into(coeff, s_sum);
}

__global__ void __launch_bounds__(kAffThreads) affine_step_kernel(const DevAffineBlock* __restrict__ blocks, long long* __restrict__ coeff)
{
  __shared__ unsigned long long s_sum[32];
  const DevAffineBlock b = blocks[blockIdx.x];
  long long*           c = coeff + (size_t) blockIdx.x * 49;
  for (int i = threadIdx.x; i < 49; i += kAffThreads) c[i] = 0;
  __syncthreads();
  if (b.sixParam) affine_step_block<true>(b, c, s_sum);
  else affine_step_block<false>(b, c, s_sum);
}

}   // namespace

cudaError_t launch_affine_sobel(const int16_t* pred, int stride, int w, int h, int vertical, int* deriv, cudaStream_t st)
{
  affine_sobel_kernel<<<(w * h + kAffThreads - 1) / kAffThreads, kAffThreads, 0, st>>>(pred, stride, w, h, vertical, deriv);
  return cudaGetLastError();
}

cudaError_t launch_affine_equal_coeff(const int16_t* residue, int residueStride, const int* d0, const int* d1, int derivStride, int w, int h,
                                      int sixParam, long long* coeff, cudaStream_t st)
{
  if (sixParam) affine_equal_coeff_kernel<true><<<1, kAffThreads, 0, st>>>(residue, residueStride, d0, d1, derivStride, w, h, coeff);
  else affine_equal_coeff_kernel<false><<<1, kAffThreads, 0, st>>>(residue, residueStride, d0, d1, derivStride, w, h, coeff);
  return cudaGetLastError();
}

cudaError_t launch_affine_step(const DevAffineBlock* dBlocks, int n, long long* dCoeff, cudaStream_t st)
{
  affine_step_kernel<<<n, kAffThreads, 0, st>>>(dBlocks, dCoeff);
  return cudaGetLastError();
}

}   // namespace vtmme
