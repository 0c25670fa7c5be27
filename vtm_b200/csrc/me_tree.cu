// Batched per-CTU integer full search (vtmme_search_frames): the reuse-minimal form of
// InterSearch::xPatternSearch (EncoderLib/InterSearch.cpp:3566-3608) over a quad-tree of square CUs.
//
// For one displacement d the SAD of a 16x16 CU is the sum of the SADs of its four 8x8 children at the same
// d (xGetSAD, CommonLib/RdCost.cpp:493-528 with subShift 0), and so on up to 128x128.  So:
//
//   me_tree_sad_kernel   one CTA per 32x32 region: stages the region's original samples (widened to 32 bit)
//                        and the reference window (16 bit) in shared memory, computes the 8x8 SADs with
//                        VABSDIFF.U32 (|a-b|+c, one instruction per pixel-candidate), sums them to 16x16
//                        and 32x32 in registers, keeps the per-CU argmin of SAD + lambda*bits (first minimum
//                        in raster order, like the reference) and writes the 32x32 SAD surface.
//   me_tree_upper_kernel one CTA per CTU slice: sums 32x32 surfaces into the 64x64 and 128x128 CUs.
//
// Each CU has its own window (xSetSearchRange around its own predictor); a region computes the bounding
// box of the windows of the 23 CUs that contain it or lie in it, and every candidate is tested against the
// window of the CU it is evaluated for.
#include <cstdio>
#include <cstdlib>

#include "me_kernels.h"

namespace vtmme {

namespace {

#ifndef VTMME_TREE_THREADS
#define VTMME_TREE_THREADS 256
#define VTMME_TREE_MINBLOCKS 3
#endif
constexpr int kTreeMaxThreads = VTMME_TREE_THREADS;
constexpr int kTreeMinBlocks  = VTMME_TREE_MINBLOCKS;   // CTAs per SM the register allocation must allow
#ifndef VTMME_ROW_UNROLL
#define VTMME_ROW_UNROLL 4
#endif
constexpr int kRowUnroll = VTMME_ROW_UNROLL;   // rows of an 8x8 block per loop trip in the hot loop
constexpr int kSlots          = 23;   // 16 8x8 + 4 16x16 + 1 32x32 (+ the 64x64 and 128x128 ancestors, window only)
constexpr int kCheckSlots     = 21;   // CUs this kernel keeps an argmin for

struct CuInfo
{
  short l, r, t, b;    // window, integer pel, inclusive
  short pqx, pqy;      // predictor, quarter-pel
  int   idx;           // index in the CU order, -1 if the CU does not exist
};

// shared-memory layout (bytes); the tables after kOffLut are sized from TreeParams::maxGx / maxRows
constexpr int kOffOrg  = 0;                       // uint32 [32][32]   original samples widened to 32 bit
constexpr int kOffBest = 4096;                    // u64 [24]          per-CU best (cost, position)
constexpr int kOffCu   = kOffBest + 24 * 8;       // CuInfo [24]
constexpr int kOffMisc = kOffCu + 24 * 16;        // int [8]
constexpr int kOffLut  = kOffMisc + 32;           // uint32 [512]      lambda * bits, entries >= 256 = "outside the window"
constexpr int kOffTab  = kOffLut + 2048;          // u8 bitsX[21][maxGx*8], u8 minBitsX[21][gxPad], u8 bitsY[21][rowsPad], then the window
constexpr uint32_t kLutInvalid = 0x3fffffffu;
static_assert(sizeof(CuInfo) == 16, "CuInfo layout");
static_assert(kOffTab % 16 == 0, "alignment");

__host__ __device__ inline int tree_gx_pad(int maxGx) { return (maxGx + 3) & ~3; }
__host__ __device__ inline int tree_rows_pad(int maxRows) { return (maxRows + 3) & ~3; }
__host__ __device__ inline int tree_off_ref(int maxGx, int maxRows)
{
  return (kOffTab + kCheckSlots * (maxGx * 8 + tree_gx_pad(maxGx) + tree_rows_pad(maxRows)) + 15) & ~15;
}

// Argmin update for 8 candidates (one 8-wide displacement group, one row) of one CU.
//   fast reject: cost >= SAD + lambda*(min bitsX of the group + bitsY of the row) =: SAD + lb
//   survivors:   exact cost SAD + lut[bitsX(dx) + bitsY(dy)] (tables hold 255 outside the CU's window, lut[>=256] is
//                "never"), atomicMin on the shared (cost, position) key = first minimum in raster order.
// The candidate with the smallest SAD goes first: on a cold threshold it tightens it for the other seven.
__device__ __forceinline__ void check8(const uint32_t (&a)[8], uint32_t lb, const uint8_t* bx8, uint32_t by,
                                       const uint32_t* lut, unsigned long long* best, int dx0, int dy)
{
  uint32_t       thr = reinterpret_cast<volatile uint2*>(best)->y;
  const uint32_t m   = min(min(min(a[0], a[1]), min(a[2], a[3])), min(min(a[4], a[5]), min(a[6], a[7])));
  if (m + lb <= thr)
  {
    const uint2 bw = *reinterpret_cast<const uint2*>(bx8);
    uint32_t    c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = a[k] + lut[(((k < 4 ? bw.x : bw.y) >> (8 * (k & 3))) & 0xffu) + by];
    int      kmin = 0;
    uint32_t cmin = c[0];
#pragma unroll
    for (int k = 7; k >= 0; k--)
      if (a[k] == m)
      {
        kmin = k;
        cmin = c[k];
      }
    if (cmin <= thr)
    {
      atomicMin(best, make_key(cmin, dx0 + kmin, dy));
      thr = reinterpret_cast<volatile uint2*>(best)->y;
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k != kmin && c[k] <= thr) atomicMin(best, make_key(c[k], dx0 + k, dy));
  }
}

// 64 pixel-candidates: 8 original samples of one row against 8 displacements.  The first 8-NFP displacements
// use VABSDIFF.U32 on the ALU pipe; the last NFP use two FADDs on the FP32 pipe: the 10-bit samples and the
// partial sums are treated as denormal floats (bit pattern == integer below 2^24), for which FADD is exact
// fixed-point arithmetic, so both halves produce the same integers.
// Packed FP32 lanes (sm_100 sub.f32x2 / add.f32x2 -> FADD2): NFP codes 12 / 13 / 14 put the displacement lanes {4,6} / {2,4,6} /
// {0,2,4,6} on FADD2, two PIXELS of one displacement per instruction: {o[i], o[i+1]} - {px[i+k], px[i+k+1]} (aligned register
// pairs because i and k are even), then acc2 += |d2| (negation and magnitude are operand modifiers).  The two halves of a
// packed accumulator are the partial sums over the even and the odd pixels; the caller adds them when the block is done.
template <int NFP>
__host__ __device__ constexpr int fp2_mask()
{
  return NFP == 12 ? 0x50 : NFP == 13 ? 0x54 : NFP == 14 ? 0x55 : 0;
}
__device__ __forceinline__ unsigned long long f2_pack(uint32_t lo, uint32_t hi)
{
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long f2_absdiff_acc(unsigned long long acc, unsigned long long o2, unsigned long long p2)
{
  unsigned long long d;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(o2), "l"(p2));
  float d0, d1;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
  unsigned long long ad;
  asm("mov.b64 %0, {%1,%2};" : "=l"(ad) : "f"(fabsf(d0)), "f"(fabsf(d1)));
  asm("add.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(ad));
  return acc;
}

template <int NFP>
__device__ __forceinline__ void sad_row(uint32_t (&a)[8], uint32_t (&ah)[8], const uint32_t (&o)[8], const uint32_t (&px)[16])
{
  constexpr int M2 = fp2_mask<NFP>();
  constexpr int NF = NFP >= 10 ? 0 : NFP;
#pragma unroll
  for (int k = 0; k < 8; k++)
  {
    if ((M2 >> k) & 1)
    {
      unsigned long long acc = f2_pack(a[k], ah[k]);
#pragma unroll
      for (int i = 0; i < 8; i += 2) acc = f2_absdiff_acc(acc, f2_pack(o[i], o[i + 1]), f2_pack(px[i + k], px[i + k + 1]));
      asm("mov.b64 {%0,%1}, %2;" : "=r"(a[k]), "=r"(ah[k]) : "l"(acc));
    }
    else if (k >= 8 - NF)
    {
      float acc = __uint_as_float(a[k]);
#pragma unroll
      for (int i = 0; i < 8; i++)
        acc = __fadd_rn(acc, fabsf(__fsub_rn(__uint_as_float(o[i]), __uint_as_float(px[i + k]))));
      a[k] = __float_as_uint(acc);
    }
    else
    {
#pragma unroll
      for (int i = 0; i < 8; i++) a[k] = __usad(o[i], px[i + k], a[k]);
    }
  }
}
template <int NFP>
__device__ __forceinline__ void sad_row(uint32_t (&a)[8], const uint32_t (&o)[8], const uint32_t (&px)[16])
{
  uint32_t unused[8];   // (paths without the second accumulator half run the packed codes as NFP = 2)
  sad_row<(NFP >= 10 ? 2 : NFP)>(a, unused, o, px);
}
// folds the halves of the packed lanes (exact: integers below 2^24 as denormal / small floats)
template <int NFP>
__device__ __forceinline__ void sad_fold(uint32_t (&a)[8], const uint32_t (&ah)[8])
{
#pragma unroll
  for (int k = 0; k < 8; k++)
    if ((fp2_mask<NFP>() >> k) & 1) a[k] = __float_as_uint(__fadd_rn(__uint_as_float(a[k]), __uint_as_float(ah[k])));
}

// The staged reference window holds two samples per 32-bit word in 12-bit fields (lo | hi << 12), so that a word
// stays below 2^24: seen as a float it is then a denormal (value = bits * 2^-149) and the FP32 pipe can split it
// exactly — w * 2^-12 rounds to hi (lo < 2048), and w - hi * 4096 = lo.
__device__ __forceinline__ uint32_t pack12(uint32_t x)   // x = lo | hi << 16, both < 1024
{
  return (x & 0xfffu) | ((x >> 4) & 0x3ff000u);
}

// 16 reference samples (two 128-bit shared loads) widened to 32 bit; FPU: on the FP32 pipe, else LOP3/SHF.
// W32: the window is staged one sample per 32-bit word (no unpack instructions at all, four 128-bit loads): a staged
// row of `refStride` samples holds its even 4-sample chunks in words [0, refStride/2) and the odd chunks in
// [refStride/2, refStride), so that the 16-byte pieces the lanes of a warp read are contiguous (no bank conflicts);
// p points at the even chunk of the first sample (a multiple of 8).
template <bool FPU, bool W32 = false>
__device__ __forceinline__ void load_ref16(const uint16_t* p, uint32_t (&px)[16], int refStride = 0)
{
  if (W32)
  {
    const uint32_t* q    = reinterpret_cast<const uint32_t*>(p);
    const int       half = refStride >> 1;
    const uint4     e0 = *reinterpret_cast<const uint4*>(q), o0 = *reinterpret_cast<const uint4*>(q + half);
    const uint4     e1 = *reinterpret_cast<const uint4*>(q + 4), o1 = *reinterpret_cast<const uint4*>(q + half + 4);
    px[0] = e0.x; px[1] = e0.y; px[2] = e0.z; px[3] = e0.w;
    px[4] = o0.x; px[5] = o0.y; px[6] = o0.z; px[7] = o0.w;
    px[8] = e1.x; px[9] = e1.y; px[10] = e1.z; px[11] = e1.w;
    px[12] = o1.x; px[13] = o1.y; px[14] = o1.z; px[15] = o1.w;
    return;
  }
  const uint4    w0 = *reinterpret_cast<const uint4*>(p);
  const uint4    w1 = *reinterpret_cast<const uint4*>(p + 8);
  const uint32_t w[8] = { w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w };
#pragma unroll
  for (int i = 0; i < 8; i++)
  {
    if (FPU)
    {
      const float wf = __uint_as_float(w[i]);
      const float hi = __fmul_rn(wf, 0.000244140625f);   // 2^-12
      px[2 * i + 1]  = __float_as_uint(hi);
      px[2 * i]      = __float_as_uint(__fmaf_rn(hi, -4096.0f, wf));
    }
    else
    {
      px[2 * i]     = w[i] & 0xfffu;
      px[2 * i + 1] = w[i] >> 12;
    }
  }
}

__device__ __forceinline__ void load_org8(const uint32_t* p, uint32_t (&o)[8])
{
  const uint4 o0 = *reinterpret_cast<const uint4*>(p);
  const uint4 o1 = *reinterpret_cast<const uint4*>(p + 4);
  o[0] = o0.x; o[1] = o0.y; o[2] = o0.z; o[3] = o0.w;
  o[4] = o1.x; o[5] = o1.y; o[6] = o1.z; o[7] = o1.w;
}

// SS (subShiftMode 2, DY == 1 only): CUs with H > 8 and W <= 64 (16x16 .. 64x64) use 2 * SAD(even rows)
// (RdCost.cpp:310-316, 489); accumulator set 0 then holds the even rows and set 1 the odd rows (8x8) / all rows.
template <int NFP, bool FPU, int DY, bool SS, bool W32>
__global__ void __launch_bounds__(kTreeMaxThreads, (DY == 2 || SS) ? 2 : kTreeMinBlocks) me_tree_sad_kernel(TreeParams p)
{
  constexpr int NACC = SS ? 2 : DY;
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t*           s_org  = reinterpret_cast<uint32_t*>(smem + kOffOrg);
  unsigned long long* s_best = reinterpret_cast<unsigned long long*>(smem + kOffBest);
  CuInfo*             s_cu   = reinterpret_cast<CuInfo*>(smem + kOffCu);
  int*                s_misc = reinterpret_cast<int*>(smem + kOffMisc);
  uint32_t*           s_lut  = reinterpret_cast<uint32_t*>(smem + kOffLut);
  const int           gxPad = tree_gx_pad(p.maxGx), rowsPad = tree_rows_pad(p.maxRows);
  const int           bxStride = p.maxGx * 8;
  uint8_t*            s_bx    = smem + kOffTab;
  uint8_t*            s_minbx = s_bx + kCheckSlots * bxStride;
  uint8_t*            s_by    = s_minbx + kCheckSlots * gxPad;
  uint16_t*           s_ref   = reinterpret_cast<uint16_t*>(smem + tree_off_ref(p.maxGx, p.maxRows));

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int region = blockIdx.x;
  const int pair   = blockIdx.y;
  const int rx = region % p.g.nRegX, ry = region / p.g.nRegX;
  const int x0 = rx * 32, y0 = ry * 32;
  const int nCU = p.g.off[5];
  const double lambda   = p.lambda;
  const int    imvShift = p.imvShift;

  // ---- 1. windows of the 23 CUs touching this region ------------------------------------------------
  if (tid < kSlots)
  {
    int level, cx, cy;
    if (tid < 16)
    {
      const int q = tid >> 2, s = tid & 3;
      level = 0;
      cx    = rx * 4 + (q & 1) * 2 + (s & 1);
      cy    = ry * 4 + (q >> 1) * 2 + (s >> 1);
    }
    else if (tid < 20)
    {
      const int q = tid - 16;
      level = 1;
      cx    = rx * 2 + (q & 1);
      cy    = ry * 2 + (q >> 1);
    }
    else
    {
      level = tid - 18;   // 2, 3, 4
      cx    = rx >> (level - 2);
      cy    = ry >> (level - 2);
    }
    CuInfo ci;
    ci.idx = -1;
    ci.l = ci.t = 32767;
    ci.r = ci.b = -32768;
    ci.pqx = ci.pqy = 0;
    if (cx < p.g.nx[level] && cy < p.g.ny[level])
    {
      ci.idx = p.g.off[level] + cy * p.g.nx[level] + cx;
      short2 pr = make_short2(0, 0);
      if (p.predQ) pr = p.predQ[(size_t) pair * nCU + ci.idx];
      const int    size = 8 << level;
      const Window w    = search_window(pr.x, pr.y, cx * size, cy * size, p.g.picW, p.g.picH, p.ctu, p.sr);
      ci.l   = (short) w.l;
      ci.r   = (short) w.r;
      ci.t   = (short) w.t;
      ci.b   = (short) w.b;
      ci.pqx = pr.x;
      ci.pqy = pr.y;
    }
    s_cu[tid]   = ci;
    s_best[tid] = ~0ull;
  }
  for (int i = tid; i < 512; i += nthr) s_lut[i] = i < 256 ? mv_cost(lambda, (uint32_t) i) : kLutInvalid;
  __syncthreads();
  if (tid == 0)
  {
    int wl = 32767, wr = -32768, wt = 32767, wb = -32768, mask8 = 0;
    for (int s = 0; s < kSlots; s++)
      if (s_cu[s].idx >= 0)
      {
        wl = min(wl, (int) s_cu[s].l);
        wr = max(wr, (int) s_cu[s].r);
        wt = min(wt, (int) s_cu[s].t);
        wb = max(wb, (int) s_cu[s].b);
        if (s < 16) mask8 |= 1 << s;
      }
    const int wl8 = wl & ~7;
    s_misc[0] = wl8;
    s_misc[1] = wt;
    s_misc[2] = (wr - wl8 + 8) >> 3;   // ngx
    s_misc[3] = wb - wt + 1;           // nrows
    s_misc[4] = mask8;
    s_misc[6] = wr - wl8 - (s_misc[2] - 1) * 8 + 1;   // valid displacements in the last 8-wide group (1..8)
  }
  __syncthreads();
  const int wl8 = s_misc[0], wt = s_misc[1], ngx = s_misc[2], nrows = s_misc[3], mask8 = s_misc[4];
  if (mask8 == 0) return;   // no CU of the set lies in this (partial) region
  if (ngx > p.maxGx || nrows > p.maxRows)
  {
    if (tid == 0) atomicExch(p.errFlag, 1);   // predictor spread larger than the context was sized for
    return;
  }
  const bool writeSurf = s_cu[20].idx >= 0 && (s_cu[21].idx >= 0 || s_cu[22].idx >= 0);
  if (tid == 0) p.regInfo[(size_t) pair * p.g.nRegX * p.g.nRegY + region] = make_int4(wl8, wt, ngx, nrows);

  // ---- 2. rate tables of the 21 CUs: min bitsX per 8-wide displacement group, bitsY per row (255 = outside) ----
  for (int i = tid; i < kCheckSlots * ngx; i += nthr)
  {
    const int    slot = i / ngx, g = i - slot * ngx;
    const CuInfo ci   = s_cu[slot];
    uint32_t     mb   = 255, lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
      const int dx = wl8 + g * 8 + k;
      uint32_t  b  = 255;
      if (ci.idx >= 0 && dx >= ci.l && dx <= ci.r) b = eg_bits((dx * 4 - ci.pqx) >> imvShift);
      mb = min(mb, b);
      if (k < 4) lo |= b << (8 * k);
      else hi |= b << (8 * (k - 4));
    }
    *reinterpret_cast<uint2*>(s_bx + slot * bxStride + g * 8) = make_uint2(lo, hi);
    s_minbx[slot * gxPad + g] = (uint8_t) mb;
  }
  for (int i = tid; i < kCheckSlots * nrows; i += nthr)
  {
    const int    slot = i / nrows, r = i - slot * nrows;
    const CuInfo ci   = s_cu[slot];
    const int    dy   = wt + r;
    uint32_t     b    = 255;
    if (ci.idx >= 0 && dy >= ci.t && dy <= ci.b) b = eg_bits((dy * 4 - ci.pqy) >> imvShift);
    s_by[slot * rowsPad + r] = (uint8_t) b;
  }

  // ---- 3. original samples of the region, widened to 32 bit -------------------------------------------
  const DevPic cur = p.cur[pair];
  const DevPic ref = p.ref[pair];
  for (int i = tid; i < 1024; i += nthr)
  {
    const int y = i >> 5, x = i & 31;
    s_org[i]    = (uint32_t) (uint16_t) cur.origin[(size_t) (y0 + y) * cur.stride + x0 + x];
  }

  const int refStride = ngx * 8 + 32;   // samples per staged row
  const int rs        = W32 ? 2 * refStride : refStride;   // the same in uint16 units of s_ref (W32: one word per sample)
  uint32_t* surf = p.surf + ((size_t) pair * p.g.nRegX * p.g.nRegY + region) * p.surfCap;
  uint32_t* surfEven = SS ? p.surfEven + ((size_t) pair * p.g.nRegX * p.g.nRegY + region) * p.surfCap : nullptr;

  // ---- 4. bands of displacement rows ----------------------------------------------------------------
  for (int band0 = blockIdx.z * p.bandRows; band0 < nrows; band0 += gridDim.z * p.bandRows)
  {
    const int bh = min(p.bandRows, nrows - band0);
    const int tileRows = (bh + DY - 1) / DY;
    __syncthreads();   // previous band fully consumed (and s_org / tables written)
    if (tid == 0) s_misc[5] = 0;   // tile counter of this band
    {
      const int      vecPerRow = refStride >> 3;
      const int      nvec      = (tileRows * DY + 31) * vecPerRow;
      const int16_t* src       = ref.origin + (ptrdiff_t) (y0 + wt + band0) * ref.stride + (x0 + wl8);
      for (int i = tid; i < nvec; i += nthr)
      {
        const int r = i / vecPerRow, c = i - r * vecPerRow;
        uint4 v = *reinterpret_cast<const uint4*>(src + (ptrdiff_t) r * ref.stride + c * 8);
        if (W32)
        {
          uint32_t* row = reinterpret_cast<uint32_t*>(s_ref) + r * refStride;
          *reinterpret_cast<uint4*>(row + c * 4) = make_uint4(v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16);
          *reinterpret_cast<uint4*>(row + (refStride >> 1) + c * 4) = make_uint4(v.z & 0xffffu, v.z >> 16, v.w & 0xffffu, v.w >> 16);
          continue;
        }
        v.x = pack12(v.x);
        v.y = pack12(v.y);
        v.z = pack12(v.z);
        v.w = pack12(v.w);
        *reinterpret_cast<uint4*>(s_ref + r * refStride + c * 8) = v;
      }
    }
    __syncthreads();

    // Tiles (8 displacements x DY rows) are handed out to warps 32 at a time from a shared counter, so that all
    // warps of the CTA finish within one tile of each other whatever the window size is.
    const int ntiles    = ngx * tileRows;
    const int lastValid = s_misc[6];
    const int ngxFull   = (DY == 1 && lastValid <= 4 && ngx > 1) ? ngx - 1 : ngx;
    const int nFull     = ngxFull * tileRows;
    for (;;)
    {
      int tbase = 0;
      if ((tid & 31) == 0) tbase = atomicAdd(&s_misc[5], 32);
      tbase = __shfl_sync(0xffffffffu, tbase, 0);
      if (tbase >= ntiles) break;
      const int t = tbase + (tid & 31);
      if (t >= ntiles) continue;
      // a last group with few valid displacements (129 = 16 * 8 + 1 for SR = 64) is handled by a narrow path; its
      // tiles are numbered after all full tiles so that whole warps take one path
      int  tr, gx;
      bool narrow = false;
      if (t < nFull)
      {
        tr = t / ngxFull;
        gx = t - tr * ngxFull;
      }
      else
      {
        tr     = t - nFull;
        gx     = ngx - 1;
        narrow = true;
      }
      const int       row0 = band0 + tr * DY;          // displacement row index of this tile's first row
      const int       dx0 = wl8 + gx * 8;
      const bool      row1ok = DY == 2 && (tr * DY + 1 < bh);
      const uint16_t* refTile = s_ref + tr * DY * rs + gx * 8;
      uint32_t        a32[NACC][8];
#pragma unroll
      for (int d = 0; d < NACC; d++)
#pragma unroll
        for (int k = 0; k < 8; k++) a32[d][k] = 0;

      for (int q = 0; q < 4; q++)
      {
        uint32_t a16[NACC][8];
#pragma unroll
        for (int d = 0; d < NACC; d++)
#pragma unroll
          for (int k = 0; k < 8; k++) a16[d][k] = 0;
        if ((mask8 >> (q * 4)) & 15)
        {
          for (int s = 0; s < 4; s++)
          {
            const int slot = q * 4 + s;
            if (!((mask8 >> slot) & 1)) continue;
            const int bx = (q & 1) * 16 + (s & 1) * 8, by = (q >> 1) * 16 + (s >> 1) * 8;
            const uint32_t* orgRow = s_org + by * 32 + bx;
            const uint16_t* refRow = refTile + by * rs + bx;
            uint32_t        a8[NACC][8];
#pragma unroll
            for (int d = 0; d < NACC; d++)
#pragma unroll
              for (int k = 0; k < 8; k++) a8[d][k] = 0;
            if (DY == 1 && narrow)
            {
              // only the first lastValid (<= 4) displacements exist: the others start from a value no minimum picks
#pragma unroll
              for (int k = 1; k < 8; k++)
                if (k >= lastValid) a8[0][k] = 1u << 24;
              if (SS)
              {
#pragma unroll
                for (int k = 1; k < 8; k++)
                  if (k >= lastValid) a8[NACC - 1][k] = 1u << 24;
              }
              const uint32_t* op = orgRow;
              const uint16_t* rp = refRow;
#pragma unroll 2
              for (int r = 0; r < 8; r++)
              {
                uint32_t o[8], px[16];
                load_org8(op, o);
                load_ref16<false, W32>(rp, px, refStride);
#pragma unroll
                for (int k = 0; k < 4; k++)
                  if (k < lastValid)
                  {
#pragma unroll
                    for (int i = 0; i < 8; i++) a8[SS ? (r & 1) : 0][k] = __usad(o[i], px[i + k], a8[SS ? (r & 1) : 0][k]);
                  }
                op += 32;
                rp += rs;
              }
            }
            else if (DY == 1)
            {
              const uint32_t* op = orgRow;
              const uint16_t* rp = refRow;
              uint32_t        a8h[NACC][8];
#pragma unroll
              for (int d = 0; d < NACC; d++)
#pragma unroll
                for (int k = 0; k < 8; k++) a8h[d][k] = 0;
#pragma unroll kRowUnroll
              for (int r = 0; r < 8; r++)
              {
                uint32_t o[8], px[16];
                load_org8(op, o);
                load_ref16<FPU, W32>(rp, px, refStride);
                sad_row<NFP>(a8[SS ? (r & 1) : 0], a8h[SS ? (r & 1) : 0], o, px);
                op += 32;
                rp += rs;
              }
#pragma unroll
              for (int d = 0; d < NACC; d++) sad_fold<NFP>(a8[d], a8h[d]);
            }
            else
            {
              // reference row jr serves original row jr at displacement row 0 and original row jr-1 at row 1;
              // kept as a rolled loop (uniform branches) so that the body stays inside the instruction cache
#pragma unroll 1
              for (int jr = 0; jr < 9; jr++)
              {
                uint32_t o[8], px[16];
                load_ref16<FPU, W32>(refRow + jr * rs, px, refStride);
                if (jr < 8)
                {
                  load_org8(orgRow + jr * 32, o);
                  sad_row<NFP>(a8[0], o, px);
                }
                if (jr > 0)
                {
                  load_org8(orgRow + (jr - 1) * 32, o);
                  sad_row<NFP>(a8[DY - 1], o, px);
                }
              }
            }
            if (SS)
            {
              // 8x8 CUs use every row; the parents need the even rows (x2 later) and, for 128x128, all rows
#pragma unroll
              for (int k = 0; k < 8; k++) a8[1][k] += a8[0][k];
              const uint32_t by = s_by[slot * rowsPad + row0];
              check8(a8[1], s_lut[s_minbx[slot * gxPad + gx] + by], s_bx + slot * bxStride + gx * 8, by, s_lut, &s_best[slot],
                     dx0, wt + row0);
#pragma unroll
              for (int k = 0; k < 8; k++)
              {
                a16[0][k] += a8[0][k];
                a16[1][k] += a8[1][k];
              }
            }
            else
            {
#pragma unroll
              for (int d = 0; d < DY; d++)
              {
                if (d == 0 || row1ok)
                {
                  const uint32_t by = s_by[slot * rowsPad + row0 + d];
                  check8(a8[d], s_lut[s_minbx[slot * gxPad + gx] + by], s_bx + slot * bxStride + gx * 8, by, s_lut,
                         &s_best[slot], dx0, wt + row0 + d);
                }
#pragma unroll
                for (int k = 0; k < 8; k++) a16[d][k] += a8[d][k];
              }
            }
          }
          if (s_cu[16 + q].idx >= 0)
          {
            if (SS)
            {
              uint32_t v[8];
#pragma unroll
              for (int k = 0; k < 8; k++) v[k] = a16[0][k] << 1;
              const uint32_t by = s_by[(16 + q) * rowsPad + row0];
              check8(v, s_lut[s_minbx[(16 + q) * gxPad + gx] + by], s_bx + (16 + q) * bxStride + gx * 8, by, s_lut,
                     &s_best[16 + q], dx0, wt + row0);
            }
            else
            {
#pragma unroll
              for (int d = 0; d < DY; d++)
                if (d == 0 || row1ok)
                {
                  const uint32_t by = s_by[(16 + q) * rowsPad + row0 + d];
                  check8(a16[d], s_lut[s_minbx[(16 + q) * gxPad + gx] + by], s_bx + (16 + q) * bxStride + gx * 8, by, s_lut,
                         &s_best[16 + q], dx0, wt + row0 + d);
                }
            }
          }
        }
#pragma unroll
        for (int d = 0; d < NACC; d++)
#pragma unroll
          for (int k = 0; k < 8; k++) a32[d][k] += a16[d][k];
      }
      if (SS)
      {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = a32[0][k] << 1;
        if (s_cu[20].idx >= 0)
        {
          const uint32_t by = s_by[20 * rowsPad + row0];
          check8(v, s_lut[s_minbx[20 * gxPad + gx] + by], s_bx + 20 * bxStride + gx * 8, by, s_lut, &s_best[20], dx0, wt + row0);
        }
        if (writeSurf)
        {
          const size_t so = (size_t) row0 * (ngx * 8) + gx * 8;
          uint4* de = reinterpret_cast<uint4*>(surfEven + so);
          de[0]     = make_uint4(v[0], v[1], v[2], v[3]);
          de[1]     = make_uint4(v[4], v[5], v[6], v[7]);
          uint4* df = reinterpret_cast<uint4*>(surf + so);
          df[0]     = make_uint4(a32[1][0], a32[1][1], a32[1][2], a32[1][3]);
          df[1]     = make_uint4(a32[1][4], a32[1][5], a32[1][6], a32[1][7]);
        }
        continue;
      }
#pragma unroll
      for (int d = 0; d < DY; d++)
      {
        if (d == 1 && !row1ok) continue;
        if (s_cu[20].idx >= 0)
        {
          const uint32_t by = s_by[20 * rowsPad + row0 + d];
          check8(a32[d], s_lut[s_minbx[20 * gxPad + gx] + by], s_bx + 20 * bxStride + gx * 8, by, s_lut, &s_best[20], dx0,
                 wt + row0 + d);
        }
        if (writeSurf)
        {
          uint4* dst = reinterpret_cast<uint4*>(surf + (size_t) (row0 + d) * (ngx * 8) + gx * 8);
          dst[0]     = make_uint4(a32[d][0], a32[d][1], a32[d][2], a32[d][3]);
          dst[1]     = make_uint4(a32[d][4], a32[d][5], a32[d][6], a32[d][7]);
        }
      }
    }
  }
  __syncthreads();
  if (tid < kCheckSlots && s_cu[tid].idx >= 0 && s_best[tid] != ~0ull)
    atomicMin(p.keys + (size_t) pair * nCU + s_cu[tid].idx, s_best[tid]);
}

// ---- 64x64 and 128x128 levels from the 32x32 surfaces ----------------------------------------------------
constexpr int kUpperThreads = 256;

// TWO: subShiftMode 2 (a second, even-row surface feeds the 64x64 level).
template <bool TWO>
__global__ void __launch_bounds__(kUpperThreads) me_tree_upper_kernel(TreeParams p)
{
  extern __shared__ __align__(16) unsigned char usmem[];
  __shared__ CuInfo          s_cu[5];
  __shared__ int4            s_reg[16];
  __shared__ const uint32_t* s_surf[16];
  __shared__ const uint32_t* s_surfE[16];   // even-row surfaces (subShiftMode 2), else the same as s_surf
  __shared__ int             s_box[4];
  __shared__ uint32_t        s_lut[512];

  const int tid  = threadIdx.x;
  const int ctu  = blockIdx.x;
  const int pair = blockIdx.y;
  const int cx = ctu % p.g.nCtuX, cy = ctu / p.g.nCtuX;
  const int nCU  = p.g.off[5];
  const int nReg = p.g.nRegX * p.g.nRegY;

  if (tid < 5)
  {
    const int level = tid < 4 ? 3 : 4;
    const int ux = tid < 4 ? cx * 2 + (tid & 1) : cx, uy = tid < 4 ? cy * 2 + (tid >> 1) : cy;
    CuInfo    ci;
    ci.idx = -1;
    ci.l = ci.t = 32767;
    ci.r = ci.b = -32768;
    ci.pqx = ci.pqy = 0;
    if (ux < p.g.nx[level] && uy < p.g.ny[level])
    {
      ci.idx = p.g.off[level] + uy * p.g.nx[level] + ux;
      short2 pr = make_short2(0, 0);
      if (p.predQ) pr = p.predQ[(size_t) pair * nCU + ci.idx];
      const int    size = 8 << level;
      const Window w    = search_window(pr.x, pr.y, ux * size, uy * size, p.g.picW, p.g.picH, p.ctu, p.sr);
      ci.l   = (short) w.l;
      ci.r   = (short) w.r;
      ci.t   = (short) w.t;
      ci.b   = (short) w.b;
      ci.pqx = pr.x;
      ci.pqy = pr.y;
    }
    s_cu[tid] = ci;
  }
  if (tid >= 32 && tid < 48)
  {
    const int i = tid - 32;
    const int rx = cx * 4 + (i & 3), ry = cy * 4 + (i >> 2);
    int4      info = make_int4(0, 0, 0, 0);
    const uint32_t *sp = nullptr, *se = nullptr;
    if (rx < p.g.nx[2] && ry < p.g.ny[2])
    {
      const size_t r = (size_t) pair * nReg + ry * p.g.nRegX + rx;
      info = p.regInfo[r];
      sp   = p.surf + r * p.surfCap;
      se   = TWO ? p.surfEven + r * p.surfCap : sp;
    }
    s_reg[i]   = info;
    s_surf[i]  = sp;
    s_surfE[i] = se;
  }
  for (int i = tid; i < 512; i += kUpperThreads) s_lut[i] = i < 256 ? mv_cost(p.lambda, (uint32_t) i) : kLutInvalid;
  __syncthreads();
  if (tid == 0)
  {
    int l = 32767, r = -32768, t = 32767, b = -32768;
    for (int s = 0; s < 5; s++)
      if (s_cu[s].idx >= 0)
      {
        l = min(l, (int) s_cu[s].l);
        r = max(r, (int) s_cu[s].r);
        t = min(t, (int) s_cu[s].t);
        b = max(b, (int) s_cu[s].b);
      }
    s_box[0] = l;
    s_box[1] = r;
    s_box[2] = t;
    s_box[3] = b;
  }
  __syncthreads();
  const int bl = s_box[0], br = s_box[1], bt = s_box[2], bb = s_box[3];
  if (br < bl) return;   // no 64x64 / 128x128 CU in this CTU
  const int bw = br - bl + 1, bhgt = bb - bt + 1;
  // rate tables of the five CUs over the bounding box (255 = outside the CU's window)
  uint8_t* s_bx = usmem;                  // [5][bw]
  uint8_t* s_by = usmem + 5 * bw;         // [5][bhgt]
  for (int i = tid; i < 5 * bw; i += kUpperThreads)
  {
    const int    s = i / bw, dx = bl + (i - s * bw);
    const CuInfo ci = s_cu[s];
    s_bx[i] = (ci.idx >= 0 && dx >= ci.l && dx <= ci.r) ? (uint8_t) eg_bits((dx * 4 - ci.pqx) >> p.imvShift) : 255;
  }
  for (int i = tid; i < 5 * bhgt; i += kUpperThreads)
  {
    const int    s = i / bhgt, dy = bt + (i - s * bhgt);
    const CuInfo ci = s_cu[s];
    s_by[i] = (ci.idx >= 0 && dy >= ci.t && dy <= ci.b) ? (uint8_t) eg_bits((dy * 4 - ci.pqy) >> p.imvShift) : 255;
  }
  __syncthreads();

  unsigned long long best[5];
#pragma unroll
  for (int s = 0; s < 5; s++) best[s] = ~0ull;

  // displacements of the box, flattened; consecutive threads read consecutive dx of the surfaces
  const int total = bw * bhgt;
  for (int i = blockIdx.z * kUpperThreads + tid; i < total; i += gridDim.z * kUpperThreads)
  {
    const int yi = i / bw, xi = i - yi * bw;
    const int dx = bl + xi, dy = bt + yi;
    uint32_t  bits[5];
    bool      any = false;
#pragma unroll
    for (int s = 0; s < 5; s++)
    {
      bits[s] = (uint32_t) s_bx[s * bw + xi] + (uint32_t) s_by[s * bhgt + yi];
      any |= bits[s] < 256;
    }
    if (!any) continue;
    // s64: what the 64x64 CUs compare (even rows x2 in subShiftMode 2); s64f: all rows, summed into the 128x128 CU.
    // All (up to 16, in subShiftMode 2 32) surface reads of a displacement are issued before any is used — the kernel is
    // bound by HBM latency x bytes in flight; a surface this displacement lies outside of reads its first element instead.
    uint32_t f[16], e[16];
#pragma unroll
    for (int ri = 0; ri < 16; ri++)
    {
      const int4      info = s_reg[ri];
      const uint32_t* sp   = s_surf[ri];
      const int       j    = ((ri >> 3) << 1) | ((ri >> 1) & 1);   // the 64x64 CU this region belongs to
      const bool      need = sp != nullptr && (bits[j] < 256 || bits[4] < 256);
      const size_t    o    = need ? (size_t) (dy - info.y) * (info.z * 8) + (dx - info.x) : 0;
      const uint32_t* pf   = need ? sp : p.surf;
      f[ri] = __ldg(pf + o);
      if (TWO) e[ri] = __ldg((need ? s_surfE[ri] : p.surf) + o);
      if (!need) f[ri] = 0;
      if (TWO && !need) e[ri] = 0;
    }
    uint32_t   s64[4], s64f[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
      s64[j] = s64f[j] = 0;
#pragma unroll
      for (int k = 0; k < 4; k++)
      {
        const int ri = ((j >> 1) * 2 + (k >> 1)) * 4 + (j & 1) * 2 + (k & 1);
        s64f[j] += f[ri];
        if (TWO) s64[j] += e[ri];
      }
    }
#pragma unroll
    for (int s = 0; s < 5; s++)
    {
      const uint32_t sad  = s < 4 ? (TWO ? s64[s] : s64f[s]) : s64f[0] + s64f[1] + s64f[2] + s64f[3];
      const uint32_t cost = sad + s_lut[bits[s]];   // >= 0x3fffffff outside the window
      const unsigned long long k = make_key(cost, dx, dy);
      if (cost < kLutInvalid && k < best[s]) best[s] = k;
    }
  }
#pragma unroll
  for (int s = 0; s < 5; s++)
  {
    unsigned long long k = best[s];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1)
    {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, m);
      k                          = o < k ? o : k;
    }
    if ((tid & 31) == 0 && k != ~0ull && s_cu[s].idx >= 0) atomicMin(p.keys + (size_t) pair * nCU + s_cu[s].idx, k);
  }
}

}   // namespace

// Tuning knobs (development): VTMME_TREE_VARIANT="nfp,fpu,dy,threads,w32,bandRows"
static void tree_variant(int& nfp, int& fpu, int& dy, int& threads, int& w32, int& band)
{
  static int v[6] = { -1, 0, 0, 0, 0, 0 };
  if (v[0] < 0)
  {
    v[0] = 2; v[1] = 1; v[2] = 1; v[3] = 256; v[4] = 1; v[5] = 0;
    if (const char* e = getenv("VTMME_TREE_VARIANT")) sscanf(e, "%d,%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]);
  }
  nfp = v[0]; fpu = v[1]; dy = v[2]; threads = v[3]; w32 = v[4]; band = v[5];
}

bool tree_window_w32()
{
  int nfp, fpu, dy, threads, w32, band;
  tree_variant(nfp, fpu, dy, threads, w32, band);
  return w32 != 0 && dy == 1;
}

int tree_band_override()
{
  int nfp, fpu, dy, threads, w32, band;
  tree_variant(nfp, fpu, dy, threads, w32, band);
  return band;
}

// Displacement rows staged per band: the largest even split of the window whose staged rows, next to the tables, leave
// room for 3 CTAs per SM (2 for the row-sub-sampling instantiation, which is register-limited to 2 anyway).
int tree_pick_band_rows(int maxGx, int maxRows, bool subSampling)
{
  if (tree_band_override() > 0) return (tree_band_override() + 1) & ~1;
  const long budget   = (subSampling ? 112L : 75L) * 1024 - tree_off_ref(maxGx, maxRows);
  const long rowBytes = (long) (maxGx * 8 + 32) * (tree_window_w32() ? 4 : 2);
  long       fit      = budget / rowBytes - 33;
  if (fit < 16) fit = 16;
  const int bands = (int) ((maxRows + fit - 1) / fit);
  return ((maxRows + bands - 1) / bands + 1) & ~1;
}

size_t tree_sad_smem_bytes(int maxGx, int maxRows, int bandRows)
{
  return (size_t) tree_off_ref(maxGx, maxRows) +
         (size_t) (bandRows + 1 + 31) * (size_t) (maxGx * 8 + 32) * (tree_window_w32() ? 4 : 2);
}

template <int NFP, bool FPU, int DY, bool SS, bool W32 = false>
static cudaError_t launch_tree_sad_t(const TreeParams& p, int nPairs, int threads, size_t smem, cudaStream_t st)
{
  static SmemOptIn optIn;
  if (cudaError_t e = optIn.ensure(me_tree_sad_kernel<NFP, FPU, DY, SS, W32>, smem)) return e;
  dim3 grid(p.g.nRegX * p.g.nRegY, nPairs, 1);
  me_tree_sad_kernel<NFP, FPU, DY, SS, W32><<<grid, threads, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_tree_sad(const TreeParams& p, int nPairs, cudaStream_t st)
{
  int nfp, fpu, dy, threads, w32, band;
  tree_variant(nfp, fpu, dy, threads, w32, band);
  w32 = tree_window_w32();
  const size_t smem = tree_sad_smem_bytes(p.maxGx, p.maxRows, p.bandRows);
  if (threads <= 0)
  {
    // even out the tile loop: interior regions have (2*sr+1 -> groups) x rows tiles
    const int ngx = (2 * p.sr + 1 + 7) / 8, rows = (2 * p.sr + 1 + dy - 1) / dy;
    const int tiles = ngx * rows, iters = (tiles + kTreeMaxThreads - 1) / kTreeMaxThreads;
    threads = ((tiles + iters - 1) / iters + 127) & ~127;   // whole multiples of 4 warps: one per SM sub-partition
    if (threads > kTreeMaxThreads) threads = kTreeMaxThreads;
    if (threads < 128) threads = 128;
  }
  if (p.subShiftMode == 2)   // row sub-sampling: one instantiation (1 displacement row per tile)
  {
    const int ngx = (2 * p.sr + 1 + 7) / 8, tiles = ngx * (2 * p.sr + 1);
    int       thr = ((tiles + (tiles + 255) / 256 - 1) / ((tiles + 255) / 256) + 127) & ~127;
    thr           = thr > kTreeMaxThreads ? kTreeMaxThreads : (thr < 128 ? 128 : thr);
    if (w32) return launch_tree_sad_t<2, true, 1, true, true>(p, nPairs, thr, smem, st);
    return launch_tree_sad_t<2, true, 1, true>(p, nPairs, thr, smem, st);
  }
  if (w32)
  {
    if (nfp == 0) return launch_tree_sad_t<0, true, 1, false, true>(p, nPairs, threads, smem, st);
    if (nfp == 1) return launch_tree_sad_t<1, true, 1, false, true>(p, nPairs, threads, smem, st);
    if (nfp == 2) return launch_tree_sad_t<2, true, 1, false, true>(p, nPairs, threads, smem, st);
    if (nfp == 3) return launch_tree_sad_t<3, true, 1, false, true>(p, nPairs, threads, smem, st);
    if (nfp == 12) return launch_tree_sad_t<12, true, 1, false, true>(p, nPairs, threads, smem, st);
    if (nfp == 13) return launch_tree_sad_t<13, true, 1, false, true>(p, nPairs, threads, smem, st);
    if (nfp == 14) return launch_tree_sad_t<14, true, 1, false, true>(p, nPairs, threads, smem, st);
    return cudaErrorInvalidValue;
  }
#define VTMME_TREE_CASE(N, F, D) \
  if (nfp == N && fpu == F && dy == D) return launch_tree_sad_t<N, F != 0, D, false>(p, nPairs, threads, smem, st);
  VTMME_TREE_CASE(0, 0, 1) VTMME_TREE_CASE(0, 1, 1) VTMME_TREE_CASE(1, 1, 1) VTMME_TREE_CASE(2, 1, 1) VTMME_TREE_CASE(3, 1, 1)
  VTMME_TREE_CASE(0, 0, 2) VTMME_TREE_CASE(0, 1, 2) VTMME_TREE_CASE(1, 1, 2) VTMME_TREE_CASE(2, 1, 2) VTMME_TREE_CASE(3, 1, 2)
  VTMME_TREE_CASE(2, 0, 2) VTMME_TREE_CASE(2, 0, 1)
#undef VTMME_TREE_CASE
  return cudaErrorInvalidValue;
}

cudaError_t launch_tree_upper(const TreeParams& p, int nPairs, cudaStream_t st)
{
  dim3 grid(p.g.nCtuX * p.g.nCtuY, nPairs, 4);
  // dynamic shared memory: rate tables over the bounding box of the CTU's five windows
  const size_t smem = (size_t) 5 * 2 * (p.maxGx * 8 + 8 + p.maxRows + 8);
  if (p.subShiftMode == 2)
    me_tree_upper_kernel<true><<<grid, kUpperThreads, smem, st>>>(p);
  else
    me_tree_upper_kernel<false><<<grid, kUpperThreads, smem, st>>>(p);
  return cudaGetLastError();
}

}   // namespace vtmme

// ---- development microbenchmark: the 8x8-block SAD inner loop alone (no staging, no argmin, no tails) ----------
namespace vtmme {
namespace {
template <int NFP, bool FPU, int DY, bool W32 = false>
__global__ void __launch_bounds__(kTreeMaxThreads, DY == 2 ? 2 : 3) sad_block_bench_kernel(uint32_t* out, int iters)
{
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t* s_org = reinterpret_cast<uint32_t*>(smem);
  uint16_t* s_ref = reinterpret_cast<uint16_t*>(smem + 4096);
  constexpr int refStride = 17 * 8 + 32;
  constexpr int rs        = W32 ? 2 * refStride : refStride;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_org[i] = (i * 2654435761u) >> 22;
  if (W32)
    for (int i = threadIdx.x; i < 48 * refStride; i += blockDim.x) reinterpret_cast<uint32_t*>(s_ref)[i] = ((i * 40503u) >> 6) & 0x3ffu;
  else
    for (int i = threadIdx.x; i < 48 * refStride / 2; i += blockDim.x)
      reinterpret_cast<uint32_t*>(s_ref)[i] = pack12((((i * 40503u) >> 6) & 0x3ffu) | ((((i * 9973u) >> 5) & 0x3ffu) << 16));
  __syncthreads();
  uint32_t acc[DY][8];
#pragma unroll
  for (int d = 0; d < DY; d++)
#pragma unroll
    for (int k = 0; k < 8; k++) acc[d][k] = 0;
  const int gx = threadIdx.x % 17, tr = (threadIdx.x / 17) % 6;
  const uint16_t* refTile = s_ref + tr * DY * rs + gx * 8;
  for (int it = 0; it < iters; it++)
  {
    for (int blk = 0; blk < 16; blk++)
    {
      const int bx = (blk & 3) * 8, by = (blk >> 2) * 8;
      const uint32_t* orgRow = s_org + by * 32 + bx;
      const uint16_t* refRow = refTile + by * rs + bx;
      uint32_t a8[DY][8];
#pragma unroll
      for (int d = 0; d < DY; d++)
#pragma unroll
        for (int k = 0; k < 8; k++) a8[d][k] = 0;
      if (DY == 1)
      {
        uint32_t a8h[8];
#pragma unroll
        for (int k = 0; k < 8; k++) a8h[k] = 0;
#pragma unroll kRowUnroll
        for (int r = 0; r < 8; r++)
        {
          uint32_t o[8], px[16];
          load_org8(orgRow + r * 32, o);
          load_ref16<FPU, W32>(refRow + r * rs, px, refStride);
          sad_row<NFP>(a8[0], a8h, o, px);
        }
        sad_fold<NFP>(a8[0], a8h);
      }
      else
      {
#pragma unroll 1
        for (int jr = 0; jr < 9; jr++)
        {
          uint32_t o[8], px[16];
          load_ref16<FPU>(refRow + jr * refStride, px);
          if (jr < 8)
          {
            load_org8(orgRow + jr * 32, o);
            sad_row<NFP>(a8[0], o, px);
          }
          if (jr > 0)
          {
            load_org8(orgRow + (jr - 1) * 32, o);
            sad_row<NFP>(a8[DY - 1], o, px);
          }
        }
      }
#pragma unroll
      for (int d = 0; d < DY; d++)
#pragma unroll
        for (int k = 0; k < 8; k++) acc[d][k] += a8[d][k];
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int d = 0; d < DY; d++)
#pragma unroll
    for (int k = 0; k < 8; k++) r += acc[d][k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NFP, bool FPU, int DY, bool W32 = false>
double run_sad_block_bench(int threads, int ctasPerSm, int iters, int sms, uint32_t* dout, cudaStream_t st)
{
  const size_t smem = 4096 + 48 * (17 * 8 + 32) * (W32 ? 4 : 2);
  cudaFuncSetAttribute(sad_block_bench_kernel<NFP, FPU, DY, W32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  sad_block_bench_kernel<NFP, FPU, DY, W32><<<sms * ctasPerSm, threads, smem, st>>>(dout, 2);
  cudaEventRecord(e0, st);
  sad_block_bench_kernel<NFP, FPU, DY, W32><<<sms * ctasPerSm, threads, smem, st>>>(dout, iters);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // pixel-candidates per second
  return (double) sms * ctasPerSm * threads * (double) iters * 16.0 * 64.0 * 8.0 * DY / (ms * 1e-3);
}
}   // namespace

// returns pixel-candidates/s of the bare inner loop for (nfp, fpu, dy) at the given launch shape
double sad_block_bench(int nfp, int fpu, int dy, int threads, int ctasPerSm, int iters)
{
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint32_t* dout = nullptr;
  cudaMalloc(&dout, (size_t) sms * ctasPerSm * threads * 4);
  cudaStream_t st;
  cudaStreamCreate(&st);
  double r = -1;
#define VTMME_BB_CASE(N, F, D) if (nfp == N && fpu == F && dy == D) r = run_sad_block_bench<N, F != 0, D>(threads, ctasPerSm, iters, sms, dout, st);
  VTMME_BB_CASE(0, 0, 1) VTMME_BB_CASE(0, 1, 1) VTMME_BB_CASE(1, 1, 1) VTMME_BB_CASE(2, 1, 1) VTMME_BB_CASE(3, 1, 1)
  VTMME_BB_CASE(0, 0, 2) VTMME_BB_CASE(0, 1, 2) VTMME_BB_CASE(1, 1, 2) VTMME_BB_CASE(2, 1, 2) VTMME_BB_CASE(3, 1, 2)
  VTMME_BB_CASE(2, 0, 2) VTMME_BB_CASE(2, 0, 1)
#undef VTMME_BB_CASE
  // fpu == 2: the one-sample-per-word window of the product path (incl. the packed FADD2 lane codes 12-14)
#define VTMME_BB_W32(N) if (nfp == N && fpu == 2 && dy == 1) r = run_sad_block_bench<N, true, 1, true>(threads, ctasPerSm, iters, sms, dout, st);
  VTMME_BB_W32(0) VTMME_BB_W32(1) VTMME_BB_W32(2) VTMME_BB_W32(3) VTMME_BB_W32(12) VTMME_BB_W32(13) VTMME_BB_W32(14)
#undef VTMME_BB_W32
  cudaStreamDestroy(st);
  cudaFree(dout);
  return r;
}
}   // namespace vtmme

extern "C" double vtmme_dev_sad_block_bench(int nfp, int fpu, int dy, int threads, int ctasPerSm, int iters)
{
  return vtmme::sad_block_bench(nfp, fpu, dy, threads, ctasPerSm, iters);
}
