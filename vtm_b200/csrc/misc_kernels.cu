// Table-level kernels (the dispatch-table flavour of the reference's SIMD entries) and picture border extension.
#include <cstdlib>
#include <type_traits>

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

// ---- SATD, one THREAD per 8x8 sub-tile (blocks with min(w, h) >= 8, power-of-two sizes, 16-byte aligned rows) ----
// MODE 0: the 8x8 tiling (w == h).  MODE 1 / 2: the 16x8 / 8x16 tilings, two threads per tile (RdCost.cpp:2369-2560).  Each thread
// transforms its own 8x8 half (L or R; T or B); the last butterfly stage of the 16-point transform pairs coefficient i of the two
// halves, and |L_i + R_i| + |L_i - R_i| = 2 max(|L_i|, |R_i|): the threads swap 32 coefficients (the first half of the tile
// takes i < 32, the second i >= 32), so a tile costs 32 shuffles per thread and no sample is loaded twice.  The DC pair keeps the
// two terms apart (|L_0 + R_0| >> 2 is the scaled DC of the tile).  The tile value (int)(sum / sqrt(128) * 2) is formed from the
// two threads' sums in FP64.  The 2-D transform of a sub-tile is 6 x 64 register butterflies, all integer, exact for any 16-bit input.

// sign-extended low / high 16-bit half of a word: PRMT with the sign-replicate bit of the selector (which __byte_perm masks off)
__device__ __forceinline__ int sext_lo(uint32_t w)
{
  int r;
  asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(r) : "r"(w));
  return r;
}
__device__ __forceinline__ int sext_hi(uint32_t w)
{
  int r;
  asm("prmt.b32 %0, %1, 0, 0xbb32;" : "=r"(r) : "r"(w));
  return r;
}
__device__ __forceinline__ void load_diff8x8(const int16_t* __restrict__ o, int os, const int16_t* __restrict__ c, int cs, int (&d)[64])
{
#pragma unroll
  for (int r = 0; r < 8; r++)
  {
    const uint4    a = *reinterpret_cast<const uint4*>(o + (size_t) r * os), b = *reinterpret_cast<const uint4*>(c + (size_t) r * cs);
    const uint32_t aw[4] = { a.x, a.y, a.z, a.w }, bw[4] = { b.x, b.y, b.z, b.w };
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
      d[r * 8 + 2 * k]     = sext_lo(aw[k]) - sext_lo(bw[k]);
      d[r * 8 + 2 * k + 1] = sext_hi(aw[k]) - sext_hi(bw[k]);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(128, 5) satd_tile_thread_kernel(const int16_t* __restrict__ org, int orgStride, long long orgBlk,
                                                               const int16_t* __restrict__ cur, int curStride, long long curBlk,
                                                               int w, int h, int n, int log2Units, int log2Ux,
                                                               unsigned long long* out)
{
  const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;   // global unit = (block, 8x8 sub-tile)
  const long long blk = g >> log2Units;
  const int       u = (int) (g & ((1 << log2Units) - 1)), ux = u & ((1 << log2Ux) - 1), uy = u >> log2Ux;
  const bool      active = blk < n;
  uint32_t        v = 0;
  int             coef[MODE == 0 ? 1 : 64];   // the transformed half of a two-thread tile
#pragma unroll
  for (int i = 0; i < (MODE == 0 ? 1 : 64); i++) coef[i] = 0;
  if (active)
  {
    const int16_t* o = org + blk * orgBlk + (size_t) (uy * 8) * orgStride + ux * 8;
    const int16_t* c = cur + blk * curBlk + (size_t) (uy * 8) * curStride + ux * 8;
    int d[64];
    load_diff8x8(o, orgStride, c, curStride, d);
#pragma unroll
    for (int st = 0; st < 6; st++)
#pragma unroll
      for (int j = 0; j < 32; j++)
      {
        const int len = 1 << st, k = ((j >> st) << (st + 1)) | (j & (len - 1));   // j-th butterfly of the stage
        const int a = d[k], b = d[k + len];
        d[k]       = a + b;
        d[k + len] = a - b;
      }
    if (MODE == 0)
    {
      uint32_t s = 0;
#pragma unroll
      for (int i = 1; i < 64; i++) s = __sad(d[i], 0, s);
      s += (uint32_t) abs(d[0]) >> 2;
      v = (s + 2) >> 2;
    }
    else
    {
#pragma unroll
      for (int i = 0; i < 64; i++) coef[i] = d[i];
    }
  }
  if (MODE != 0)
  {
    // the two halves of a tile sit in lanes that differ in one bit: bit 0 (MODE 1) or bit log2Ux (MODE 2); inactive lanes
    // (beyond the batch) only take part in the shuffles
    const int      pd = MODE == 1 ? 1 : (1 << log2Ux);
    const int      half = MODE == 1 ? (ux & 1) : (uy & 1);
    uint32_t       s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++)
    {
      const int own = half ? coef[32 + i] : coef[i];
      const int got = __shfl_xor_sync(0xffffffffu, half ? coef[i] : coef[32 + i], pd);
      if (i == 0)
      {
        const uint32_t both = 2u * (uint32_t) max(abs(own), abs(got));                       // coefficient 32 (second half)
        const uint32_t dc   = ((uint32_t) abs(own + got) >> 2) + (uint32_t) abs(own - got);   // coefficient 0: scaled DC + its pair
        s += half ? both : dc;
      }
      else
        s += 2u * (uint32_t) max(abs(own), abs(got));
    }
    const uint32_t t = s + __shfl_xor_sync(0xffffffffu, s, pd);
    v = (half || !active) ? 0u : (uint32_t) (int) __dmul_rn(__ddiv_rn((double) (int) t, 0x1.6a09e667f3bcdp+3), 2.0);
  }
  // sum over the units of a block: segments of min(units, 32) lanes; more than 32 units -> one atomic per warp
  const int seg = log2Units < 5 ? (1 << log2Units) : 32;
  for (int m = seg >> 1; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  if (active && (threadIdx.x & (seg - 1)) == 0)
  {
    if (log2Units <= 5) out[blk] = v;
    else atomicAdd(out + blk, (unsigned long long) v);
  }
}

// ---- SATD of blocks with a 4-sample side: one THREAD per 8x4 / 4x8 / 4x4 tile (RdCost.cpp:2594-2817, 2166-2265) ----
// TW x TH samples in registers, the separable Hadamard as log2(TW * TH) stages of register butterflies, DC >> 2, the tile's
// normalisation ((int)(s / sqrt(32) * 2) in FP64; (s + 1) >> 1 for 4x4), tiles of a block summed over min(tiles, 32) lanes.
template <int TW, int TH>
__global__ void __launch_bounds__(128) satd_small_tile_kernel(const int16_t* __restrict__ org, int orgStride, long long orgBlk,
                                                              const int16_t* __restrict__ cur, int curStride, long long curBlk,
                                                              int n, int log2Tiles, int log2Tx, unsigned long long* out)
{
  constexpr int   N = TW * TH;
  const long long g = (long long) blockIdx.x * blockDim.x + threadIdx.x;   // global tile
  const long long blk = g >> log2Tiles;
  const int       t = (int) (g & ((1 << log2Tiles) - 1)), tx = t & ((1 << log2Tx) - 1), ty = t >> log2Tx;
  const bool      active = blk < n;
  uint32_t        v = 0;
  if (active)
  {
    const int16_t* o = org + blk * orgBlk + (size_t) (ty * TH) * orgStride + tx * TW;
    const int16_t* c = cur + blk * curBlk + (size_t) (ty * TH) * curStride + tx * TW;
    int d[N];
#pragma unroll
    for (int r = 0; r < TH; r++)
    {
      uint32_t aw[TW / 2], bw[TW / 2];
      if (TW == 8)
      {
        const uint4 a = *reinterpret_cast<const uint4*>(o + (size_t) r * orgStride), b = *reinterpret_cast<const uint4*>(c + (size_t) r * curStride);
        aw[0] = a.x; aw[1] = a.y; aw[TW / 2 - 2] = a.z; aw[TW / 2 - 1] = a.w;
        bw[0] = b.x; bw[1] = b.y; bw[TW / 2 - 2] = b.z; bw[TW / 2 - 1] = b.w;
      }
      else
      {
        const uint2 a = *reinterpret_cast<const uint2*>(o + (size_t) r * orgStride), b = *reinterpret_cast<const uint2*>(c + (size_t) r * curStride);
        aw[0] = a.x; aw[1] = a.y;
        bw[0] = b.x; bw[1] = b.y;
      }
#pragma unroll
      for (int k = 0; k < TW / 2; k++)
      {
        d[r * TW + 2 * k]     = sext_lo(aw[k]) - sext_lo(bw[k]);
        d[r * TW + 2 * k + 1] = sext_hi(aw[k]) - sext_hi(bw[k]);
      }
    }
    constexpr int STAGES = N == 32 ? 5 : 4;
#pragma unroll
    for (int st = 0; st < STAGES; st++)
#pragma unroll
      for (int j = 0; j < N / 2; j++)
      {
        const int len = 1 << st, k = ((j >> st) << (st + 1)) | (j & (len - 1));
        const int a = d[k], b = d[k + len];
        d[k]       = a + b;
        d[k + len] = a - b;
      }
    uint32_t s = 0;
#pragma unroll
    for (int i = 1; i < N; i++) s = __sad(d[i], 0, s);
    s += (uint32_t) abs(d[0]) >> 2;
    if (N == 16) v = (s + 1) >> 1;
    else v = (uint32_t) (int) __dmul_rn(__ddiv_rn((double) (int) s, 0x1.6a09e667f3bcdp+2), 2.0);   // / sqrt(32) * 2
  }
  const int seg = log2Tiles < 5 ? (1 << log2Tiles) : 32;
  for (int m = seg >> 1; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  if (active && (threadIdx.x & (seg - 1)) == 0)
  {
    if (log2Tiles <= 5) out[blk] = v;
    else atomicAdd(out + blk, (unsigned long long) v);
  }
}

// ---- interpolation: InterpolationFilter::filter<N,...> / filterCopy (InterpolationFilter.cpp:397-656) with the
//      public dispatch of filterHor/filterVer (:749-895).  interp_batch_kernel is the generic form (any width, copies, 2-tap
//      filters); the fast paths for 4- and 8-tap filters follow it. ----
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

// Fast path of the same filters (4 or 8 taps, even width): TWO horizontally adjacent outputs per thread, no shared memory, no
// barrier — the taps are read straight from global memory / L1 as 32-bit words (neighbouring threads re-read the same lines out
// of L1), so the kernel is a pure stream with many independent loads in flight per thread.
//   * the taps are 2-way dot products IDP.2A (two 16-bit samples x two 8-bit coefficients);
//   * horizontal: outputs x, x+1 need the nine (five) samples s[x .. x+TAPS]: four (two) aligned words and one half-word.  If
//     s[x] is word-aligned, output x takes the coefficient pairs (c0,c1)(c2,c3).. and output x+1 the pairs shifted by one tap
//     (0,c0)(c1,c2)..(c7,0) over the same words; if it is not, the roles swap — no sample is ever shifted or unpacked;
//   * vertical: the words of rows y+k and y+k+1 are interleaved into (row k, row k+1) pairs per column (PRMT), then the same dot
//     products; a column pair at an odd sample address is assembled from two half-word loads;
//   * the two results leave as one 32-bit store when the destination allows it.
// Only samples of the block and its tap halo are read.  Arithmetic is InterpolationFilter::filter's (sum in 32 bits, + offset,
// >> shift, clip if last, store as Pel).
__device__ __forceinline__ uint32_t pack_taps(const int16_t* c, int i0)   // bytes: c[i0], c[i0+1], c[i0+2], c[i0+3] (0 outside 0..7)
{
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 4; k++)
  {
    const int i = i0 + k;
    const int v = (i >= 0 && i < 8) ? (int) c[i] : 0;
    r |= ((uint32_t) v & 0xffu) << (8 * k);
  }
  return r;
}

// x / d as one IMAD.HI with m = floor((2^32 - 1) / d) + 1 (d == 1: m wraps to 0).  m * d = 2^32 + e with 0 <= e <= d, and the
// quotient is exact while x * e < 2^32 — always for a power of two (e = 0); the host checks the largest dividend.
struct FastDiv
{
  uint32_t d, m;
  __host__ FastDiv(int div = 1) : d((uint32_t) div), m((uint32_t) (0xffffffffu / (uint32_t) div) + 1u) {}
  __host__ bool exact_below(unsigned long long maxX) const
  {
    if (d == 1) return maxX <= 0xffffffffull;
    const unsigned long long e = (unsigned long long) m * d - (1ull << 32);
    return maxX <= 0xffffffffull && (e == 0 || maxX * e < (1ull << 32));
  }
  __device__ __forceinline__ uint32_t div(uint32_t x) const { return d == 1 ? x : __umulhi(x, m); }
};

__device__ __forceinline__ uint32_t ld_u16(const int16_t* p) { return (uint32_t) * reinterpret_cast<const uint16_t*>(p); }

template <int TAPS, bool VERT>
__global__ void __launch_bounds__(kInterpThreads) interp_pairs_kernel(InterpArgs a, uint32_t units, int dstWords, FastDiv dHalfW, FastDiv dH)
{
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
  constexpr int  before = TAPS / 2 - 1;
  const uint32_t cA0 = pack_taps(a.coeff, 0), cA1 = pack_taps(a.coeff, 4);                                  // (c0 c1 c2 c3) (c4 c5 c6 c7)
  const uint32_t cS0 = pack_taps(a.coeff, -1), cS1 = pack_taps(a.coeff, 3), cS2 = pack_taps(a.coeff, 7);   // (0 c0 c1 c2) (c3 c4 c5 c6) (c7 0 0 0)
  // persistent grid: a CTA that lives for 256 units is bound by the CTA launch rate, not by memory
#pragma unroll 2
  for (uint32_t u = blockIdx.x * kInterpThreads + threadIdx.x; u < units; u += gridDim.x * kInterpThreads)
  {
  const uint32_t by = dHalfW.div(u), xp = u - by * dHalfW.d, b = dH.div(by), y = by - b * dH.d;
  int s0 = offset, s1 = offset;
  if (!VERT)
  {
    const int16_t* s = a.src + (long long) b * a.srcBlk + (ptrdiff_t) y * a.srcStride + (2 * (int) xp - before);   // first tap of output x
    const bool     odd = (reinterpret_cast<uintptr_t>(s) & 2) != 0;
    // aligned: W[0..TAPS/2-1] = words (s[0],s[1]).., W[TAPS/2] = s[TAPS] in the low half
    // odd:     W[0] = s[0] in the HIGH half, W[1..TAPS/2] = words (s[1],s[2])..
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(odd ? s + 1 : s);
    int             W[TAPS / 2 + 1];
    if (odd)
    {
      W[0] = (int) (ld_u16(s) << 16);
#pragma unroll
      for (int k = 0; k < TAPS / 2; k++) W[k + 1] = (int) wp[k];
    }
    else
    {
#pragma unroll
      for (int k = 0; k < TAPS / 2; k++) W[k] = (int) wp[k];
      W[TAPS / 2] = (int) ld_u16(s + TAPS);
    }
    int sa = offset, ss = offset;   // sa: aligned coefficient pairs over 4 (2) words; ss: shifted pairs over 5 (3) words
    const int* A = odd ? W + 1 : W;
    sa = __dp2a_lo(A[0], (int) cA0, sa);
    sa = __dp2a_hi(A[1], (int) cA0, sa);
    ss = __dp2a_lo(W[0], (int) cS0, ss);
    ss = __dp2a_hi(W[1], (int) cS0, ss);
    ss = __dp2a_lo(W[2], (int) cS1, ss);
    if (TAPS == 8)
    {
      sa = __dp2a_lo(A[2], (int) cA1, sa);
      sa = __dp2a_hi(A[3], (int) cA1, sa);
      ss = __dp2a_hi(W[3], (int) cS1, ss);
      ss = __dp2a_lo(W[4], (int) cS2, ss);
    }
    s0 = odd ? ss : sa;
    s1 = odd ? sa : ss;
  }
  else
  {
    const int16_t* s = a.src + (long long) b * a.srcBlk + (ptrdiff_t) ((int) y - before) * a.srcStride + 2 * xp;
    uint32_t       R[TAPS];
#pragma unroll
    for (int k = 0; k < TAPS; k++)
    {
      const int16_t* q = s + (ptrdiff_t) k * a.srcStride;
      if ((reinterpret_cast<uintptr_t>(q) & 2) == 0) R[k] = *reinterpret_cast<const uint32_t*>(q);
      else R[k] = ld_u16(q) | (ld_u16(q + 1) << 16);
    }
#pragma unroll
    for (int k = 0; k < TAPS; k += 2)
    {
      const int lo = (int) __byte_perm(R[k], R[k + 1], 0x5410), hi = (int) __byte_perm(R[k], R[k + 1], 0x7632);
      const int c  = (int) (k < 4 ? cA0 : cA1);
      if ((k & 2) == 0)
      {
        s0 = __dp2a_lo(lo, c, s0);
        s1 = __dp2a_lo(hi, c, s1);
      }
      else
      {
        s0 = __dp2a_hi(lo, c, s0);
        s1 = __dp2a_hi(hi, c, s1);
      }
    }
  }
  int v0 = (int16_t) (s0 >> shift), v1 = (int16_t) (s1 >> shift);
  if (a.isLast)
  {
    v0 = min(max(v0, 0), maxv);
    v1 = min(max(v1, 0), maxv);
  }
  int16_t* d = a.dst + (long long) b * a.dstBlk + (size_t) y * a.dstStride + 2 * xp;
  if (dstWords) *reinterpret_cast<uint32_t*>(d) = (uint32_t) (uint16_t) v0 | ((uint32_t) (uint16_t) v1 << 16);
  else
  {
    d[0] = (int16_t) v0;
    d[1] = (int16_t) v1;
  }
  }
}

// The same arithmetic with EIGHT (horizontal: 8 adjacent outputs of a row) or SIXTEEN (vertical: 2 columns x 8 rows) outputs per
// thread: the index arithmetic, the loads and — vertically — the interleaved row pairs are shared between the outputs, which is
// what makes these filters memory- instead of instruction-bound.  Horizontal needs w % 8 == 0, vertical h % 8 == 0.
// Horizontal: the 8 + TAPS - 1 samples of a thread come in as the two or three ALIGNED 16-byte chunks that hold them (LDG.128: the
// lanes of a warp read 512 contiguous bytes per instruction; with 32-bit loads the lanes were 16 bytes apart and the kernel was
// bound by the L1 data pipe, 0.71 of the HBM peak).  Where the first sample sits in its chunk (q = 0..7 samples, the same for a
// whole row) only selects which registers the dot products read: word offset q >> 1 as a compile-time offset into the loaded
// words, parity q & 1 as the swap of the two coefficient sets — a warp-uniform switch over eight instantiations, no data movement.
// A chunk is only loaded if it holds at least one sample of the block or its halo (the third one is predicated), so no access
// leaves the 16-byte granule of a needed sample.
template <int TAPS>
__global__ void __launch_bounds__(kInterpThreads) interp_hor8_kernel(InterpArgs a, uint32_t units, int dstVec, FastDiv dW8, FastDiv dH)
{
  const int hr   = max(2, 14 - a.bitDepth);
  const int maxv = (1 << a.bitDepth) - 1;
  int       shift = 6, offset;
  if (a.isLast)
  {
    shift  = 6;                      // filterHor is a first stage: first && last
    offset = 1 << (shift - 1);
  }
  else
  {
    shift -= hr;
    offset = -(8192 << shift);
  }
  constexpr int  before = TAPS / 2 - 1, NW = (8 + TAPS) / 2;   // words that hold the 8 + TAPS - 1 samples of a thread
  const uint32_t cA0 = pack_taps(a.coeff, 0), cA1 = pack_taps(a.coeff, 4);
  const uint32_t cS0 = pack_taps(a.coeff, -1), cS1 = pack_taps(a.coeff, 3), cS2 = pack_taps(a.coeff, 7);
  for (uint32_t u = blockIdx.x * kInterpThreads + threadIdx.x; u < units; u += gridDim.x * kInterpThreads)
  {
    const uint32_t by = dW8.div(u), xg = u - by * dW8.d, b = dH.div(by), y = by - b * dH.d;
    const int16_t* s = a.src + (long long) b * a.srcBlk + (ptrdiff_t) y * a.srcStride + (8 * (int) xg - before);   // first tap of output 8 xg
    const int      q = (int) ((reinterpret_cast<uintptr_t>(s) >> 1) & 7);   // sample offset of s[0] in its 16-byte chunk
    const uint4*   cp = reinterpret_cast<const uint4*>(reinterpret_cast<uintptr_t>(s) & ~(uintptr_t) 15);
    const uint4    c0 = cp[0], c1 = cp[1];
    uint4          c2 = make_uint4(0, 0, 0, 0);
    if ((q >> 1) + NW > 8) c2 = cp[2];
    const int L[12] = { (int) c0.x, (int) c0.y, (int) c0.z, (int) c0.w, (int) c1.x, (int) c1.y, (int) c1.z, (int) c1.w,
                        (int) c2.x, (int) c2.y, (int) c2.z, (int) c2.w };
    int v[8];
    // W[k] = L[WO + k].  Even parity: W[k] = (s[2k], s[2k+1]): output i even -> pairs (c0,c1).. from word i/2, odd -> shifted pairs
    // (0,c0)(c1,c2)..(c7,0) from word (i-1)/2.  Odd parity: W[k] = (s[2k-1], s[2k]): even i -> shifted pairs from word i/2, odd i ->
    // pairs from word (i+1)/2.  A half-word outside the taps always meets a zero coefficient.
    auto outputs = [&](auto woTag, auto oddTag) {
      constexpr int  WO  = decltype(woTag)::value;
      constexpr bool ODD = decltype(oddTag)::value;
#pragma unroll
      for (int i = 0; i < 8; i++)
      {
        int acc = offset;
        if (((i & 1) != 0) == ODD)
        {
          const int ia = WO + (i + 1) / 2;
          acc = __dp2a_lo(L[ia], (int) cA0, acc);
          acc = __dp2a_hi(L[ia + 1], (int) cA0, acc);
          if (TAPS == 8)
          {
            acc = __dp2a_lo(L[ia + 2], (int) cA1, acc);
            acc = __dp2a_hi(L[ia + 3], (int) cA1, acc);
          }
        }
        else
        {
          const int is = WO + i / 2;
          acc = __dp2a_lo(L[is], (int) cS0, acc);
          acc = __dp2a_hi(L[is + 1], (int) cS0, acc);
          acc = __dp2a_lo(L[is + 2], (int) cS1, acc);
          if (TAPS == 8)
          {
            acc = __dp2a_hi(L[is + 3], (int) cS1, acc);
            acc = __dp2a_lo(L[is + 4], (int) cS2, acc);
          }
        }
        int r = acc >> shift;
        if (a.isLast) r = __vimin_s32_relu(r, maxv);   // ClipPel of the 32-bit value, then the store truncates to a Pel
        v[i] = r;
      }
    };
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    using I3 = std::integral_constant<int, 3>;
    switch (q)
    {
      case 0: outputs(I0{}, std::false_type{}); break;
      case 1: outputs(I0{}, std::true_type{}); break;
      case 2: outputs(I1{}, std::false_type{}); break;
      case 3: outputs(I1{}, std::true_type{}); break;
      case 4: outputs(I2{}, std::false_type{}); break;
      case 5: outputs(I2{}, std::true_type{}); break;
      case 6: outputs(I3{}, std::false_type{}); break;
      default: outputs(I3{}, std::true_type{}); break;
    }
    int16_t* d = a.dst + (long long) b * a.dstBlk + (size_t) y * a.dstStride + 8 * xg;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = __byte_perm((uint32_t) v[2 * k], (uint32_t) v[2 * k + 1], 0x5410);
    if (dstVec == 2) *reinterpret_cast<uint4*>(d) = make_uint4(o[0], o[1], o[2], o[3]);
    else if (dstVec == 1)
    {
#pragma unroll
      for (int k = 0; k < 4; k++) reinterpret_cast<uint32_t*>(d)[k] = o[k];
    }
    else
    {
#pragma unroll
      for (int i = 0; i < 8; i++) d[i] = (int16_t) v[i];
    }
  }
}

// Vertical: requires word-aligned column pairs (source base 4-byte aligned, even strides) and source / destination spans below
// 2 GB: thread offsets are 32-bit, added once to uniform bases (per-row 64-bit index arithmetic cost a quarter of the instructions).
template <int TAPS>
__global__ void __launch_bounds__(kInterpThreads) interp_ver8_kernel(InterpArgs a, uint32_t units, FastDiv dHalfW, FastDiv dH8)
{
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
  constexpr int  before = TAPS / 2 - 1, NR = 8 + TAPS - 1;
  const uint32_t cA0 = pack_taps(a.coeff, 0), cA1 = pack_taps(a.coeff, 4);
  const char*    srcBase = reinterpret_cast<const char*>(a.src - (ptrdiff_t) before * a.srcStride);   // first tap row of output row 0
  char*          dstBase = reinterpret_cast<char*>(a.dst);
  const uint32_t srcRowBytes = 2u * (uint32_t) a.srcStride, dstRowBytes = 2u * (uint32_t) a.dstStride;
  const uint32_t srcBlkBytes = 2u * (uint32_t) a.srcBlk, dstBlkBytes = 2u * (uint32_t) a.dstBlk;
  for (uint32_t u = blockIdx.x * kInterpThreads + threadIdx.x; u < units; u += gridDim.x * kInterpThreads)
  {
    // unit = (block, group of 8 rows, column pair); column pairs fastest: a warp reads 128 contiguous bytes per row
    const uint32_t by = dHalfW.div(u), xp = u - by * dHalfW.d, b = dH8.div(by), yg = by - b * dH8.d;
    const uint32_t so = b * srcBlkBytes + 8u * yg * srcRowBytes + 4u * xp;
    uint32_t       R[NR];
    const char*    sp = srcBase + so;   // per-thread base; the row offsets k * srcRowBytes are uniform
#pragma unroll
    for (int k = 0; k < NR; k++) R[k] = *reinterpret_cast<const uint32_t*>(sp + (size_t) (k * srcRowBytes));
    // rows k and k+1 interleaved, column x (lo) / x+1 (hi)
    int lo[NR - 1], hi[NR - 1];
#pragma unroll
    for (int k = 0; k < NR - 1; k++)
    {
      lo[k] = (int) __byte_perm(R[k], R[k + 1], 0x5410);
      hi[k] = (int) __byte_perm(R[k], R[k + 1], 0x7632);
    }
    char* dp = dstBase + (b * dstBlkBytes + 8u * yg * dstRowBytes + 4u * xp);
#pragma unroll
    for (int r = 0; r < 8; r++)
    {
      int s0 = offset, s1 = offset;
      s0 = __dp2a_lo(lo[r], (int) cA0, s0);
      s1 = __dp2a_lo(hi[r], (int) cA0, s1);
      s0 = __dp2a_hi(lo[r + 2], (int) cA0, s0);
      s1 = __dp2a_hi(hi[r + 2], (int) cA0, s1);
      if (TAPS == 8)
      {
        s0 = __dp2a_lo(lo[r + 4], (int) cA1, s0);
        s1 = __dp2a_lo(hi[r + 4], (int) cA1, s1);
        s0 = __dp2a_hi(lo[r + 6], (int) cA1, s0);
        s1 = __dp2a_hi(hi[r + 6], (int) cA1, s1);
      }
      int v0 = s0 >> shift, v1 = s1 >> shift;
      if (a.isLast)   // ClipPel of the 32-bit value (one instruction), as InterpolationFilter::filter does before it stores a Pel
      {
        v0 = __vimin_s32_relu(v0, maxv);
        v1 = __vimin_s32_relu(v1, maxv);
      }
      *reinterpret_cast<uint32_t*>(dp + (size_t) (r * dstRowBytes)) = __byte_perm((uint32_t) v0, (uint32_t) v1, 0x5410);
    }
  }
}

// VTMME_INTERP_VARIANT=generic keeps every call on interp_batch_kernel (comparison runs)
static bool interp_generic_only()
{
  static int v = -1;
  if (v < 0)
  {
    const char* e = getenv("VTMME_INTERP_VARIANT");
    v = (e && e[0] == 'g') ? 1 : 0;
  }
  return v == 1;
}

static cudaError_t launch_interp(const InterpArgs& a, int n, cudaStream_t st)
{
  const int halo = a.copy ? 0 : a.taps - 1;
  const int sw = a.w + (a.vertical ? 0 : halo), sh = a.h + (a.vertical ? halo : 0);
  int bpc = 2048 / (a.w * a.h);
  bpc     = bpc < 1 ? 1 : (bpc > 32 ? 32 : bpc);
  bool tapsFitBytes = true;   // IDP.2A multiplies 16-bit samples by 8-bit coefficients (every VVC filter table fits)
  for (int k = 0; k < 8; k++) tapsFitBytes = tapsFitBytes && a.coeff[k] >= -128 && a.coeff[k] <= 127;
  const unsigned long long outPairs = (unsigned long long) n * a.h * (a.w / 2);
  if (!a.copy && (a.taps == 8 || a.taps == 4) && (a.w & 1) == 0 && !interp_generic_only() && tapsFitBytes && outPairs < (1ULL << 31))
  {
    const int dstWords = (reinterpret_cast<uintptr_t>(a.dst) & 3) == 0 && (a.dstStride & 1) == 0 && (a.dstBlk & 1) == 0;
    auto grid = [](uint32_t units) {
      const uint32_t ctas = (units + kInterpThreads - 1) / kInterpThreads;
      return ctas > 148u * 16u ? 148u * 16u : ctas;   // persistent: 8 CTAs of 256 threads fit an SM, two rounds
    };
    if (!a.vertical && (a.w & 7) == 0 && a.isFirst)
    {
      const int      dstVec = (reinterpret_cast<uintptr_t>(a.dst) & 15) == 0 && (a.dstStride & 7) == 0 && (a.dstBlk & 7) == 0 ? 2 : dstWords;
      const uint32_t units  = (uint32_t) n * (uint32_t) a.h * (uint32_t) (a.w / 8);
      const FastDiv  dW8(a.w / 8), dH(a.h);
      if (dW8.exact_below(units) && dH.exact_below(units / dW8.d))
      {
        if (a.taps == 8) interp_hor8_kernel<8><<<grid(units), kInterpThreads, 0, st>>>(a, units, dstVec, dW8, dH);
        else interp_hor8_kernel<4><<<grid(units), kInterpThreads, 0, st>>>(a, units, dstVec, dW8, dH);
        return cudaGetLastError();
      }
    }
    // byte spans of the batch: the vertical fast path addresses with 32-bit offsets from uniform row bases
    const unsigned long long srcSpan = 2ull * ((unsigned long long) (n - 1) * (unsigned long long) a.srcBlk + (unsigned long long) (a.h + 8) * a.srcStride + a.w);
    const unsigned long long dstSpan = 2ull * ((unsigned long long) (n - 1) * (unsigned long long) a.dstBlk + (unsigned long long) a.h * a.dstStride + a.w);
    const bool srcWords = (reinterpret_cast<uintptr_t>(a.src) & 3) == 0 && (a.srcStride & 1) == 0 && (a.srcBlk & 1) == 0;
    if (a.vertical && (a.h & 7) == 0 && dstWords && srcWords && a.srcBlk >= 0 && a.dstBlk >= 0 && a.srcStride > 0 && a.dstStride > 0 &&
        srcSpan < (1ull << 31) && dstSpan < (1ull << 31))
    {
      const uint32_t units = (uint32_t) n * (uint32_t) (a.h / 8) * (uint32_t) (a.w / 2);
      const FastDiv  dHalf(a.w / 2), dHg(a.h / 8);
      if (dHalf.exact_below(units) && dHg.exact_below(units / dHalf.d))
      {
        if (a.taps == 8) interp_ver8_kernel<8><<<grid(units), kInterpThreads, 0, st>>>(a, units, dHalf, dHg);
        else interp_ver8_kernel<4><<<grid(units), kInterpThreads, 0, st>>>(a, units, dHalf, dHg);
        return cudaGetLastError();
      }
    }
    const uint32_t units = (uint32_t) outPairs;
    const FastDiv  dHalfW(a.w / 2), dH(a.h);
    if (dHalfW.exact_below(units) && dH.exact_below(units / dHalfW.d))
    {
#define VTMME_INTERP_PAIRS(T, V) \
      { interp_pairs_kernel<T, V><<<grid(units), kInterpThreads, 0, st>>>(a, units, dstWords, dHalfW, dH); return cudaGetLastError(); }
      if (a.taps == 8 && !a.vertical) VTMME_INTERP_PAIRS(8, false)
      if (a.taps == 8 && a.vertical) VTMME_INTERP_PAIRS(8, true)
      if (a.taps == 4 && !a.vertical) VTMME_INTERP_PAIRS(4, false)
      VTMME_INTERP_PAIRS(4, true)
#undef VTMME_INTERP_PAIRS
    }
  }
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

// VTMME_SATD_VARIANT=warp keeps every call on satd_batch_kernel (one warp per block; comparison runs)
static bool satd_warp_only()
{
  static int v = -1;
  if (v < 0)
  {
    const char* e = getenv("VTMME_SATD_VARIANT");
    v = (e && e[0] == 'w') ? 1 : 0;
  }
  return v == 1;
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
  if (pow2 && w >= 8 && h >= 8 && !satd_warp_only() && aligned_layout(org, orgStride, orgBlockStride, 8) &&
      aligned_layout(cur, curStride, curBlockStride, 8))
  {
    int log2Ux = 0, log2Units = 0;
    while ((8 << log2Ux) < w) log2Ux++;
    while ((64LL << log2Units) < (long long) w * h) log2Units++;
    const long long units = (long long) n << log2Units;
    const int       ctas  = (int) ((units + 127) / 128);
    if (log2Units > 5)
      if (cudaError_t e = cudaMemsetAsync(out, 0, (size_t) n * sizeof(unsigned long long), st)) return e;
#define VTMME_SATD_TT(M) satd_tile_thread_kernel<M><<<ctas, 128, 0, st>>>(org, orgStride, orgBlockStride, cur, curStride, curBlockStride, w, h, n, log2Units, log2Ux, out)
    if (w == h) VTMME_SATD_TT(0);
    else if (w > h) VTMME_SATD_TT(1);
    else VTMME_SATD_TT(2);
#undef VTMME_SATD_TT
    return cudaGetLastError();
  }
  // blocks with a 4-sample side: 8x4 tiles (w > h = 4), 4x8 tiles (h > w = 4), the 4x4 block
  if (pow2 && (w == 4 || h == 4) && w <= 128 && h <= 128 && !satd_warp_only())
  {
    const int tw = (w > h) ? 8 : 4, th = (h > w) ? 8 : 4;
    if (aligned_layout(org, orgStride, orgBlockStride, tw) && aligned_layout(cur, curStride, curBlockStride, tw))
    {
      int log2Tx = 0, log2Tiles = 0;
      while ((tw << log2Tx) < w) log2Tx++;
      while (((long long) tw * th << log2Tiles) < (long long) w * h) log2Tiles++;
      const long long tiles = (long long) n << log2Tiles;
      const int       ctas  = (int) ((tiles + 127) / 128);
#define VTMME_SATD_ST(TW_, TH_) satd_small_tile_kernel<TW_, TH_><<<ctas, 128, 0, st>>>(org, orgStride, orgBlockStride, cur, curStride, curBlockStride, n, log2Tiles, log2Tx, out)
      if (tw == 8) VTMME_SATD_ST(8, 4);
      else if (th == 8) VTMME_SATD_ST(4, 8);
      else VTMME_SATD_ST(4, 4);
#undef VTMME_SATD_ST
      return cudaGetLastError();
    }
  }
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
