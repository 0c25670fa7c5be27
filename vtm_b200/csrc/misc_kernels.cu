// Table-level kernels (the dispatch-table flavour of the reference's SIMD entries) and picture border extension.
#include "me_kernels.h"

namespace vtmme {

namespace {

// ---- Picture::extendPicBorder (CommonLib/Picture.cpp:1050-1096): replicate edge samples into the margin ----
__global__ void extend_border_kernel(DevPic p)
{
  const int W = p.width + 2 * p.margin, H = p.height + 2 * p.margin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < W * H; i += gridDim.x * blockDim.x)
  {
    const int y = i / W - p.margin, x = i % W - p.margin;
    if (x >= 0 && x < p.width && y >= 0 && y < p.height) continue;
    const int sx = min(max(x, 0), p.width - 1), sy = min(max(y, 0), p.height - 1);
    p.origin[(ptrdiff_t) y * p.stride + x] = p.origin[(ptrdiff_t) sy * p.stride + sx];
  }
}

// Same replication, the picture area coming from a contiguous staging copy of the caller's plane (row stride
// srcStride): one pass writes the whole padded plane.  Used by the host uploads — a 2-D DMA of 1080 narrow rows is an
// order of magnitude slower than one contiguous transfer.
__global__ void scatter_extend_kernel(DevPic p, const int16_t* __restrict__ staging, int srcStride)
{
  const int W = p.width + 2 * p.margin, H = p.height + 2 * p.margin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < W * H; i += gridDim.x * blockDim.x)
  {
    const int y = i / W - p.margin, x = i % W - p.margin;
    const int sx = min(max(x, 0), p.width - 1), sy = min(max(y, 0), p.height - 1);
    p.origin[(ptrdiff_t) y * p.stride + x] = staging[(size_t) sy * srcStride + sx];
  }
}

// ---- SAD: RdCost::xGetSAD (RdCost.cpp:493-528) / xGetSAD_NxN_SIMD (x86/RdCostX86.h:341-456) ----
// Generic layout: one warp per block, scalar loads.
__global__ void __launch_bounds__(256) sad_batch_kernel(const int16_t* __restrict__ org, int orgStride, long long orgBlk,
                                                        const int16_t* __restrict__ cur, int curStride, long long curBlk,
                                                        int w, int h, int subShift, int n, unsigned long long* out)
{
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int16_t* o = org + (long long) warp * orgBlk;
  const int16_t* c = cur + (long long) warp * curBlk;
  const int rows = h >> subShift, step = 1 << subShift;
  uint32_t  s = 0;
  for (int i = lane; i < rows * w; i += 32)
  {
    const int r = (i / w) * step, x = i % w;
    s += (uint32_t) abs((int) o[(size_t) r * orgStride + x] - (int) c[(size_t) r * curStride + x]);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (lane == 0) out[warp] = (unsigned long long) s << subShift;
}

// Aligned layout (strides, block strides and bases multiples of VEC samples): VEC-sample vector loads (128-bit for
// VEC = 8, 64-bit for VEC = 4), several small blocks per warp so that every lane has a vector to load.
__device__ __forceinline__ uint32_t sad_packed(uint32_t a, uint32_t b, uint32_t acc)
{
  acc = __sad((int) (short) (a & 0xffffu), (int) (short) (b & 0xffffu), acc);
  return __sad((int) a >> 16, (int) b >> 16, acc);
}

template <int VEC>
__global__ void __launch_bounds__(256) sad_batch_vec_kernel(const int16_t* __restrict__ org, int orgStride, long long orgBlk,
                                                            const int16_t* __restrict__ cur, int curStride, long long curBlk,
                                                            int w, int h, int subShift, int n, unsigned long long* out)
{
  const int vecPerRow = w / VEC, rows = h >> subShift, vecPerBlock = vecPerRow * rows;
  const int lanesPerBlock = vecPerBlock >= 32 ? 32 : vecPerBlock;   // power of two
  const int blocksPerWarp = 32 / lanesPerBlock;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int blk = warp * blocksPerWarp + lane / lanesPerBlock, lib = lane % lanesPerBlock;
  uint32_t  s = 0;
  if (blk < n)
  {
    const int16_t* o = org + (long long) blk * orgBlk;
    const int16_t* c = cur + (long long) blk * curBlk;
    for (int v = lib; v < vecPerBlock; v += lanesPerBlock)
    {
      const int    r = (v / vecPerRow) << subShift, x = (v % vecPerRow) * VEC;
      const size_t oo = (size_t) r * orgStride + x, co = (size_t) r * curStride + x;
      if (VEC == 8)
      {
        const uint4 a = *reinterpret_cast<const uint4*>(o + oo), b = *reinterpret_cast<const uint4*>(c + co);
        s = sad_packed(a.x, b.x, s);
        s = sad_packed(a.y, b.y, s);
        s = sad_packed(a.z, b.z, s);
        s = sad_packed(a.w, b.w, s);
      }
      else
      {
        const uint2 a = *reinterpret_cast<const uint2*>(o + oo), b = *reinterpret_cast<const uint2*>(c + co);
        s = sad_packed(a.x, b.x, s);
        s = sad_packed(a.y, b.y, s);
      }
    }
  }
  for (int m = lanesPerBlock >> 1; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (blk < n && lib == 0) out[blk] = (unsigned long long) s << subShift;
}

// ---- SATD: RdCost::xGetHADs tiling (RdCost.cpp:2819-2934); one warp per block, TH lanes per tile ----
template <int TW, int TH, bool ALIGNED>
__device__ __forceinline__ uint32_t satd_block(const int16_t* o, int os, const int16_t* c, int cs, int w, int h, int lane)
{
  const int tilesX = w / TW, nTiles = tilesX * (h / TH);
  const int group = lane / TH, lit = lane % TH, nGroups = 32 / TH;
  uint32_t  sum = 0;
  for (int t0 = 0; t0 < nTiles; t0 += nGroups)
  {
    const int  t      = t0 + group;
    const bool active = t < nTiles;
    const int  tt     = active ? t : 0;
    const int  tx = (tt % tilesX) * TW, ty = (tt / tilesX) * TH + lit;
    int        d[TW];
    const int16_t* op = o + (size_t) ty * os + tx;
    const int16_t* cp = c + (size_t) ty * cs + tx;
    if (ALIGNED && TW >= 8)
    {
#pragma unroll
      for (int i = 0; i < TW; i += 8)
      {
        const uint4 a = *reinterpret_cast<const uint4*>(op + i), b = *reinterpret_cast<const uint4*>(cp + i);
        const uint32_t aw[4] = { a.x, a.y, a.z, a.w }, bw[4] = { b.x, b.y, b.z, b.w };
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
          d[i + 2 * k]     = (int) (short) (aw[k] & 0xffffu) - (int) (short) (bw[k] & 0xffffu);
          d[i + 2 * k + 1] = ((int) aw[k] >> 16) - ((int) bw[k] >> 16);
        }
      }
    }
    else
    {
#pragma unroll
      for (int i = 0; i < TW; i++) d[i] = (int) op[i] - (int) cp[i];
    }
    const uint32_t v = satd_tile_rows<TW, TH>(d, lit);
    if (active && lit == 0) sum += v;
  }
  return sum;
}

__device__ __forceinline__ uint32_t satd_block_2x2(const int16_t* o, int os, const int16_t* c, int cs, int w, int h, int lane)
{
  uint32_t sum = 0;   // RdCost::xCalcHADs2x2, RdCost.cpp:2140-2164
  const int tilesX = w / 2, nTiles = tilesX * (h / 2);
  for (int t = lane; t < nTiles; t += 32)
  {
    const int tx = (t % tilesX) * 2, ty = (t / tilesX) * 2;
    const int d0 = o[(size_t) ty * os + tx] - c[(size_t) ty * cs + tx], d1 = o[(size_t) ty * os + tx + 1] - c[(size_t) ty * cs + tx + 1];
    const int d2 = o[(size_t) (ty + 1) * os + tx] - c[(size_t) (ty + 1) * cs + tx];
    const int d3 = o[(size_t) (ty + 1) * os + tx + 1] - c[(size_t) (ty + 1) * cs + tx + 1];
    const int m0 = d0 + d2, m1 = d1 + d3, m2 = d0 - d2, m3 = d1 - d3;
    sum += (uint32_t) (abs(m0 + m1) >> 2) + (uint32_t) abs(m0 - m1) + (uint32_t) abs(m2 + m3) + (uint32_t) abs(m2 - m3);
  }
  return sum;
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256) satd_batch_kernel(const int16_t* __restrict__ org, int orgStride, long long orgBlk,
                                                         const int16_t* __restrict__ cur, int curStride, long long curBlk,
                                                         int w, int h, int n, unsigned long long* out)
{
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int16_t* o = org + (long long) warp * orgBlk;
  const int16_t* c = cur + (long long) warp * curBlk;
  int tw, th;
  satd_tiling(w, h, tw, th);
  uint32_t s;
  if (tw == 16) s = satd_block<16, 8, ALIGNED>(o, orgStride, c, curStride, w, h, lane);
  else if (th == 16) s = satd_block<8, 16, ALIGNED>(o, orgStride, c, curStride, w, h, lane);
  else if (tw == 8 && th == 8) s = satd_block<8, 8, ALIGNED>(o, orgStride, c, curStride, w, h, lane);
  else if (tw == 8 && th == 4) s = satd_block<8, 4, ALIGNED>(o, orgStride, c, curStride, w, h, lane);
  else if (tw == 4 && th == 8) s = satd_block<4, 8, false>(o, orgStride, c, curStride, w, h, lane);
  else if (tw == 4) s = satd_block<4, 4, false>(o, orgStride, c, curStride, w, h, lane);
  else s = satd_block_2x2(o, orgStride, c, curStride, w, h, lane);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (lane == 0) out[warp] = s;
}

// ---- interpolation: InterpolationFilter::filter<N,...> / filterCopy (InterpolationFilter.cpp:397-656) with the
//      public dispatch of filterHor/filterVer (:749-895).  One CTA per block. ----
struct InterpArgs
{
  const int16_t* src;
  int16_t*       dst;
  long long      srcBlk, dstBlk;
  int            srcStride, dstStride, w, h, taps, vertical, isFirst, isLast, bitDepth, copy;
  int16_t        coeff[8];
};

// The source block with its tap halo is staged in shared memory with row-contiguous loads, every output then reads its
// taps from there; a CTA handles `bpc` consecutive blocks so that small blocks still fill it.
constexpr int kInterpThreads = 256;

__global__ void __launch_bounds__(kInterpThreads) interp_batch_kernel(InterpArgs a, int n, int bpc)
{
  extern __shared__ __align__(16) int16_t s_src[];
  const int hr   = max(2, 14 - a.bitDepth);
  const int maxv = (1 << a.bitDepth) - 1;
  int       shift = 6, offset;
  if (a.isLast)
  {
    shift += a.isFirst ? 0 : hr;
    offset = 1 << (shift - 1);
    offset += a.isFirst ? 0 : (8192 << 6);
  }
  else
  {
    shift -= a.isFirst ? hr : 0;
    offset = a.isFirst ? -(8192 << shift) : 0;
  }
  const int before = a.copy ? 0 : a.taps / 2 - 1, halo = a.copy ? 0 : a.taps - 1;
  const int sw = a.w + (a.vertical ? 0 : halo), sh = a.h + (a.vertical ? halo : 0);
  const int tile = sw * sh;
  const int first = blockIdx.x * bpc, count = min(bpc, n - first);
  // stage: block b of this CTA at s_src + b * tile; one warp per source row (no per-element index arithmetic)
  const int  warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nWarps = kInterpThreads / 32;
  const bool wide = a.w >= 32;   // narrow blocks: flat indexing keeps all lanes busy
  if (!wide)
  {
    for (int i = threadIdx.x; i < count * tile; i += kInterpThreads)
    {
      const int b = i / tile, r = (i - b * tile) / sw, c = (i - b * tile) - r * sw;
      const int16_t* src = a.src + (long long) (first + b) * a.srcBlk;
      s_src[i] = a.vertical ? src[(ptrdiff_t) (r - before) * a.srcStride + c] : src[(ptrdiff_t) r * a.srcStride + (c - before)];
    }
  }
  else
  for (int row = warp; row < count * sh; row += nWarps)
  {
    const int      b = row / sh, r = row - b * sh;
    const int16_t* src = a.src + (long long) (first + b) * a.srcBlk +
                         (a.vertical ? (ptrdiff_t) (r - before) * a.srcStride : (ptrdiff_t) r * a.srcStride - before);
    int16_t* d = s_src + b * tile + r * sw;
    for (int c = lane; c < sw; c += 32) d[c] = src[c];
  }
  __syncthreads();
  const int tStride = a.vertical ? sw : 1;
  int       cf[8];
#pragma unroll
  for (int k = 0; k < 8; k++) cf[k] = a.coeff[k];
  const int outs = a.w * a.h;
  const int units = wide ? count * a.h : count * outs;   // wide: one warp per output row; narrow: one thread per output
  for (int u = wide ? warp : threadIdx.x; u < units; u += wide ? nWarps : kInterpThreads)
  {
    int b, y, x0, xStep;
    if (wide)
    {
      b = u / a.h;
      y = u - b * a.h;
      x0 = lane;
      xStep = 32;
    }
    else
    {
      b = u / outs;
      y = (u - b * outs) / a.w;
      x0 = (u - b * outs) - y * a.w;
      xStep = a.w;   // exactly one output
    }
    const int16_t* srow = s_src + b * tile + y * sw;
    int16_t*       drow = a.dst + (long long) (first + b) * a.dstBlk + (size_t) y * a.dstStride;
    for (int x = x0; x < a.w; x += xStep)
    {
      const int16_t* sp = srow + x;   // first tap of this output
      int            v;
      if (a.copy)
      {
        if (a.isFirst == a.isLast) v = sp[0];
        else if (a.isFirst) v = (int16_t) ((int16_t) (sp[0] << hr) - (int16_t) 8192);
        else
        {
          v = (int16_t) ((sp[0] + 8192 + (1 << (hr - 1))) >> hr);
          v = min(max(v, 0), maxv);
        }
      }
      else
      {
        int sum = 0;
        if (a.taps == 8)
        {
#pragma unroll
          for (int k = 0; k < 8; k++) sum += (int) sp[k * tStride] * cf[k];
        }
        else
        {
          for (int k = 0; k < a.taps; k++) sum += (int) sp[k * tStride] * cf[k];
        }
        v = (int16_t) ((sum + offset) >> shift);
        if (a.isLast) v = min(max(v, 0), maxv);
      }
      drow[x] = (int16_t) v;
    }
  }
}

static cudaError_t launch_interp(const InterpArgs& a, int n, cudaStream_t st)
{
  const int halo = a.copy ? 0 : a.taps - 1;
  const int sw = a.w + (a.vertical ? 0 : halo), sh = a.h + (a.vertical ? halo : 0);
  int bpc = 2048 / (a.w * a.h);
  bpc     = bpc < 1 ? 1 : (bpc > 32 ? 32 : bpc);
  const size_t smem = (size_t) bpc * sw * sh * sizeof(int16_t);
  static SmemOptIn optIn;
  if (cudaError_t e = optIn.ensure(interp_batch_kernel, smem, 48 * 1024)) return e;
  interp_batch_kernel<<<(n + bpc - 1) / bpc, kInterpThreads, smem, st>>>(a, n, bpc);
  return cudaGetLastError();
}

const int16_t h_luma[16][8] = {
  { 0, 0, 0, 64, 0, 0, 0, 0 },       { 0, 1, -3, 63, 4, -2, 1, 0 },     { -1, 2, -5, 62, 8, -3, 1, 0 },
  { -1, 3, -8, 60, 13, -4, 1, 0 },   { -1, 4, -10, 58, 17, -5, 1, 0 },  { -1, 4, -11, 52, 26, -8, 3, -1 },
  { -1, 3, -9, 47, 31, -10, 4, -1 }, { -1, 4, -11, 45, 34, -10, 4, -1 }, { -1, 4, -11, 40, 40, -11, 4, -1 },
  { -1, 4, -10, 34, 45, -11, 4, -1 }, { -1, 4, -10, 31, 47, -9, 3, -1 }, { -1, 3, -8, 26, 52, -11, 4, -1 },
  { 0, 1, -5, 17, 58, -10, 4, -1 },  { 0, 1, -4, 13, 60, -8, 3, -1 },   { 0, 1, -3, 8, 62, -5, 2, -1 },
  { 0, 1, -2, 4, 63, -3, 1, 0 }
};
const int16_t h_luma4x4[16][8] = {
  { 0, 0, 0, 64, 0, 0, 0, 0 },     { 0, 1, -3, 63, 4, -2, 1, 0 },   { 0, 1, -5, 62, 8, -3, 1, 0 },
  { 0, 2, -8, 60, 13, -4, 1, 0 },  { 0, 3, -10, 58, 17, -5, 1, 0 }, { 0, 3, -11, 52, 26, -8, 2, 0 },
  { 0, 2, -9, 47, 31, -10, 3, 0 }, { 0, 3, -11, 45, 34, -10, 3, 0 }, { 0, 3, -11, 40, 40, -11, 3, 0 },
  { 0, 3, -10, 34, 45, -11, 3, 0 }, { 0, 3, -10, 31, 47, -9, 2, 0 }, { 0, 2, -8, 26, 52, -11, 3, 0 },
  { 0, 1, -5, 17, 58, -10, 3, 0 }, { 0, 1, -4, 13, 60, -8, 2, 0 },  { 0, 1, -3, 8, 62, -5, 1, 0 },
  { 0, 1, -2, 4, 63, -3, 1, 0 }
};
const int16_t h_alt[8]        = { 0, 3, 9, 20, 20, 9, 3, 0 };
const int16_t h_chroma[32][4] = {
  { 0, 64, 0, 0 },    { -1, 63, 2, 0 },   { -2, 62, 4, 0 },   { -2, 60, 7, -1 },  { -2, 58, 10, -2 }, { -3, 57, 12, -2 },
  { -4, 56, 14, -2 }, { -4, 55, 15, -2 }, { -4, 54, 16, -2 }, { -5, 53, 18, -2 }, { -6, 52, 20, -2 }, { -6, 49, 24, -3 },
  { -6, 46, 28, -4 }, { -5, 44, 29, -4 }, { -4, 42, 30, -4 }, { -4, 39, 33, -4 }, { -4, 36, 36, -4 }, { -4, 33, 39, -4 },
  { -4, 30, 42, -4 }, { -4, 29, 44, -5 }, { -4, 28, 46, -6 }, { -3, 24, 49, -6 }, { -2, 20, 52, -6 }, { -2, 18, 53, -5 },
  { -2, 16, 54, -4 }, { -2, 15, 55, -4 }, { -2, 14, 56, -4 }, { -2, 12, 57, -3 }, { -2, 10, 58, -2 }, { -1, 7, 60, -2 },
  { 0, 4, 62, -2 },   { 0, 2, 63, -1 }
};

}   // namespace

cudaError_t launch_scatter_extend(DevPic pic, const int16_t* staging, int srcStride, cudaStream_t st)
{
  scatter_extend_kernel<<<592, 256, 0, st>>>(pic, staging, srcStride);
  return cudaGetLastError();
}

cudaError_t launch_extend_border(DevPic pic, cudaStream_t st)
{
  extend_border_kernel<<<592, 256, 0, st>>>(pic);
  return cudaGetLastError();
}

static bool aligned_layout(const void* p, int stride, long long blk, int vec)
{
  return (reinterpret_cast<uintptr_t>(p) % (2 * vec)) == 0 && stride % vec == 0 && blk % vec == 0;
}

cudaError_t launch_dist_batch(int kind, const int16_t* org, int orgStride, long long orgBlockStride, const int16_t* cur,
                              int curStride, long long curBlockStride, int w, int h, int subShift, int n,
                              unsigned long long* out, cudaStream_t st)
{
  const bool pow2 = (w & (w - 1)) == 0 && (h & (h - 1)) == 0;
  if (kind == 0)
  {
    const int vec = (w % 8 == 0) ? 8 : (w % 4 == 0 ? 4 : 0);
    if (vec && pow2 && aligned_layout(org, orgStride, orgBlockStride, vec) && aligned_layout(cur, curStride, curBlockStride, vec))
    {
      const int vecPerBlock = (w / vec) * (h >> subShift);
      const int blocksPerWarp = vecPerBlock >= 32 ? 1 : 32 / vecPerBlock;
      const int warps = (n + blocksPerWarp - 1) / blocksPerWarp, ctas = (warps + 7) / 8;
      if (vec == 8)
        sad_batch_vec_kernel<8><<<ctas, 256, 0, st>>>(org, orgStride, orgBlockStride, cur, curStride, curBlockStride, w, h, subShift, n, out);
      else
        sad_batch_vec_kernel<4><<<ctas, 256, 0, st>>>(org, orgStride, orgBlockStride, cur, curStride, curBlockStride, w, h, subShift, n, out);
      return cudaGetLastError();
    }
    sad_batch_kernel<<<(n + 7) / 8, 256, 0, st>>>(org, orgStride, orgBlockStride, cur, curStride, curBlockStride, w, h,
                                                  subShift, n, out);
    return cudaGetLastError();
  }
  const int blocks = (n + 7) / 8;   // 8 warps per CTA
  if (aligned_layout(org, orgStride, orgBlockStride, 8) && aligned_layout(cur, curStride, curBlockStride, 8))
    satd_batch_kernel<true><<<blocks, 256, 0, st>>>(org, orgStride, orgBlockStride, cur, curStride, curBlockStride, w, h, n, out);
  else
    satd_batch_kernel<false><<<blocks, 256, 0, st>>>(org, orgStride, orgBlockStride, cur, curStride, curBlockStride, w, h, n, out);
  return cudaGetLastError();
}

cudaError_t launch_filter_batch(int taps, int vertical, int isFirst, int isLast, int copy, const int16_t* src, int srcStride,
                                long long srcBlockStride, int16_t* dst, int dstStride, long long dstBlockStride, int w, int h,
                                const int16_t* coeff, int bitDepth, int n, cudaStream_t st)
{
  InterpArgs a;
  a.src = src;
  a.dst = dst;
  a.srcBlk = srcBlockStride;
  a.dstBlk = dstBlockStride;
  a.srcStride = srcStride;
  a.dstStride = dstStride;
  a.w = w;
  a.h = h;
  a.vertical = vertical;
  a.isFirst  = isFirst;
  a.isLast   = isLast;
  a.bitDepth = bitDepth;
  a.copy     = copy;
  a.taps     = taps;
  for (int k = 0; k < 8; k++) a.coeff[k] = (!copy && k < taps) ? coeff[k] : 0;
  return launch_interp(a, n, st);
}

cudaError_t launch_interp_batch(int comp, int vertical, const int16_t* src, int srcStride, long long srcBlockStride,
                                int16_t* dst, int dstStride, long long dstBlockStride, int w, int h, int frac, int isFirst,
                                int isLast, int bitDepth, int useAltHpel, int n, cudaStream_t st)
{
  InterpArgs a;
  a.src = src;
  a.dst = dst;
  a.srcBlk = srcBlockStride;
  a.dstBlk = dstBlockStride;
  a.srcStride = srcStride;
  a.dstStride = dstStride;
  a.w = w;
  a.h = h;
  a.vertical = vertical;
  a.isFirst  = vertical ? isFirst : 1;   // filterHor is always a first stage (InterpolationFilter.cpp:677-689)
  a.isLast   = isLast;
  a.bitDepth = bitDepth;
  a.copy     = frac == 0;
  a.taps     = comp == 0 ? 8 : 4;
  const int16_t* c;
  if (comp == 0)
  {
    // coefficient choice of filterHor / filterVer (InterpolationFilter.cpp:782-794, 865-877)
    const bool q4 = vertical ? (w == 4 && h == 4) : ((w == 4 && h == 4) || (w == 4 && h == 11));
    if (frac == 8 && useAltHpel) c = h_alt;
    else if (q4) c = h_luma4x4[frac];
    else c = h_luma[frac];
  }
  else
    c = h_chroma[frac];
  for (int k = 0; k < 8; k++) a.coeff[k] = k < a.taps ? c[k] : 0;
  return launch_interp(a, n, st);
}

}   // namespace vtmme
