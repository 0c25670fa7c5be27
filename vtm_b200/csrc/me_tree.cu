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
#include "me_kernels.h"

namespace vtmme {

namespace {

constexpr int kTreeThreads = 256;
constexpr int kSlots       = 23;   // 16 8x8 + 4 16x16 + 1 32x32 (+ the 64x64 and 128x128 ancestors, window only)

struct CuInfo
{
  short l, r, t, b;    // window, integer pel, inclusive
  short pqx, pqy;      // predictor, quarter-pel
  int   idx;           // index in the CU order, -1 if the CU does not exist
};

// shared-memory layout (bytes)
constexpr int kOffOrg  = 0;                       // uint32 [32][32]
constexpr int kOffBest = 4096;                    // u64 [24]
constexpr int kOffCu   = kOffBest + 24 * 8;       // CuInfo [24]
constexpr int kOffMisc = kOffCu + 24 * 16;        // int [8]
constexpr int kOffRef  = kOffMisc + 32;           // uint16 [rows][refStride]
static_assert(sizeof(CuInfo) == 16, "CuInfo layout");
static_assert(kOffRef % 16 == 0, "alignment");

// Candidate that passed the cheap SAD <= current-best test: exact window test, exact cost, atomic argmin.
__device__ __noinline__ void consider(const CuInfo* cu, unsigned long long* best, int dx, int dy, uint32_t sad,
                                      double lambda, int imvShift)
{
  if (dx < cu->l || dx > cu->r || dy < cu->t || dy > cu->b) return;
  const uint32_t cost = sad + mv_cost(lambda, mv_bits_q(dx * 4, dy * 4, cu->pqx, cu->pqy, imvShift));
  atomicMin(best, make_key(cost, dx, dy));
}

__device__ __forceinline__ void check8(const uint32_t (&a)[8], const CuInfo* cu, unsigned long long* best, int dx0,
                                       int dy, double lambda, int imvShift)
{
  const uint32_t thr = (uint32_t) (*reinterpret_cast<volatile unsigned long long*>(best) >> 32);
  const uint32_t m   = min(min(min(a[0], a[1]), min(a[2], a[3])), min(min(a[4], a[5]), min(a[6], a[7])));
  if (m <= thr)
  {
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (a[k] <= thr) consider(cu, best, dx0 + k, dy, a[k], lambda, imvShift);
  }
}

__global__ void __launch_bounds__(kTreeThreads, 3) me_tree_sad_kernel(TreeParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t*           s_org  = reinterpret_cast<uint32_t*>(smem + kOffOrg);
  unsigned long long* s_best = reinterpret_cast<unsigned long long*>(smem + kOffBest);
  CuInfo*             s_cu   = reinterpret_cast<CuInfo*>(smem + kOffCu);
  int*                s_misc = reinterpret_cast<int*>(smem + kOffMisc);
  uint16_t*           s_ref  = reinterpret_cast<uint16_t*>(smem + kOffRef);

  const int tid    = threadIdx.x;
  const int region = blockIdx.x;
  const int pair   = blockIdx.y;
  const int rx = region % p.g.nRegX, ry = region / p.g.nRegX;
  const int x0 = rx * 32, y0 = ry * 32;
  const int nCU = p.g.off[5];

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

  // ---- 2. original samples of the region, widened to 32 bit -------------------------------------------
  const DevPic cur = p.cur[pair];
  const DevPic ref = p.ref[pair];
  for (int i = tid; i < 1024; i += kTreeThreads)
  {
    const int y = i >> 5, x = i & 31;
    s_org[i]    = (uint32_t) (uint16_t) cur.origin[(size_t) (y0 + y) * cur.stride + x0 + x];
  }

  const int refStride = ngx * 8 + 32;   // samples per staged row
  uint32_t* surf = p.surf + ((size_t) pair * p.g.nRegX * p.g.nRegY + region) * p.surfCap;
  const double lambda   = p.lambda;
  const int    imvShift = p.imvShift;

  // ---- 3. bands of displacement rows ----------------------------------------------------------------
  for (int band0 = blockIdx.z * p.bandRows; band0 < nrows; band0 += gridDim.z * p.bandRows)
  {
    const int bh = min(p.bandRows, nrows - band0);
    __syncthreads();   // previous band fully consumed (and s_org written)
    {
      const int      vecPerRow = refStride >> 3;
      const int      nvec      = (bh + 31) * vecPerRow;
      const int16_t* src       = ref.origin + (ptrdiff_t) (y0 + wt + band0) * ref.stride + (x0 + wl8);
      for (int i = tid; i < nvec; i += kTreeThreads)
      {
        const int r = i / vecPerRow, c = i - r * vecPerRow;
        const uint4 v = *reinterpret_cast<const uint4*>(src + (ptrdiff_t) r * ref.stride + c * 8);
        *reinterpret_cast<uint4*>(s_ref + r * refStride + c * 8) = v;
      }
    }
    __syncthreads();

    const int ntiles = ngx * bh;
    for (int t = tid; t < ntiles; t += kTreeThreads)
    {
      const int       dyi = t / ngx, gx = t - dyi * ngx;
      const int       dy = wt + band0 + dyi, dx0 = wl8 + gx * 8;
      const uint16_t* refTile = s_ref + dyi * refStride + gx * 8;
      uint32_t        a32[8];
#pragma unroll
      for (int k = 0; k < 8; k++) a32[k] = 0;

      for (int q = 0; q < 4; q++)
      {
        uint32_t a16[8];
#pragma unroll
        for (int k = 0; k < 8; k++) a16[k] = 0;
        if ((mask8 >> (q * 4)) & 15)
        {
          for (int s = 0; s < 4; s++)
          {
            const int slot = q * 4 + s;
            if (!((mask8 >> slot) & 1)) continue;
            const int bx = (q & 1) * 16 + (s & 1) * 8, by = (q >> 1) * 16 + (s >> 1) * 8;
            uint32_t  a8[8];
#pragma unroll
            for (int k = 0; k < 8; k++) a8[k] = 0;
#pragma unroll
            for (int r = 0; r < 8; r++)
            {
              const uint4 o0 = *reinterpret_cast<const uint4*>(s_org + (by + r) * 32 + bx);
              const uint4 o1 = *reinterpret_cast<const uint4*>(s_org + (by + r) * 32 + bx + 4);
              const uint4 w0 = *reinterpret_cast<const uint4*>(refTile + (by + r) * refStride + bx);
              const uint4 w1 = *reinterpret_cast<const uint4*>(refTile + (by + r) * refStride + bx + 8);
              const uint32_t o[8]   = { o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w };
              const uint32_t px[16] = { w0.x & 0xffffu, w0.x >> 16, w0.y & 0xffffu, w0.y >> 16,
                                        w0.z & 0xffffu, w0.z >> 16, w0.w & 0xffffu, w0.w >> 16,
                                        w1.x & 0xffffu, w1.x >> 16, w1.y & 0xffffu, w1.y >> 16,
                                        w1.z & 0xffffu, w1.z >> 16, w1.w & 0xffffu, w1.w >> 16 };
#pragma unroll
              for (int k = 0; k < 8; k++)
#pragma unroll
                for (int i = 0; i < 8; i++) a8[k] = __usad(o[i], px[i + k], a8[k]);
            }
            check8(a8, &s_cu[slot], &s_best[slot], dx0, dy, lambda, imvShift);
#pragma unroll
            for (int k = 0; k < 8; k++) a16[k] += a8[k];
          }
          if (s_cu[16 + q].idx >= 0) check8(a16, &s_cu[16 + q], &s_best[16 + q], dx0, dy, lambda, imvShift);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) a32[k] += a16[k];
      }
      if (s_cu[20].idx >= 0) check8(a32, &s_cu[20], &s_best[20], dx0, dy, lambda, imvShift);
      if (writeSurf)
      {
        uint4* dst = reinterpret_cast<uint4*>(surf + (size_t) (band0 + dyi) * (ngx * 8) + gx * 8);
        dst[0]     = make_uint4(a32[0], a32[1], a32[2], a32[3]);
        dst[1]     = make_uint4(a32[4], a32[5], a32[6], a32[7]);
      }
    }
  }
  __syncthreads();
  if (tid < 21 && s_cu[tid].idx >= 0 && s_best[tid] != ~0ull)
    atomicMin(p.keys + (size_t) pair * nCU + s_cu[tid].idx, s_best[tid]);
}

// ---- 64x64 and 128x128 levels from the 32x32 surfaces ----------------------------------------------------
constexpr int kUpperThreads = 256;

__global__ void __launch_bounds__(kUpperThreads) me_tree_upper_kernel(TreeParams p)
{
  __shared__ CuInfo          s_cu[5];
  __shared__ int4            s_reg[16];
  __shared__ const uint32_t* s_surf[16];
  __shared__ int             s_box[4];

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
    const uint32_t* sp = nullptr;
    if (rx < p.g.nx[2] && ry < p.g.ny[2])
    {
      const size_t r = (size_t) pair * nReg + ry * p.g.nRegX + rx;
      info = p.regInfo[r];
      sp   = p.surf + r * p.surfCap;
    }
    s_reg[i]  = info;
    s_surf[i] = sp;
  }
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

  unsigned long long best[5];
#pragma unroll
  for (int s = 0; s < 5; s++) best[s] = ~0ull;

  const int total = bw * bhgt;
  for (int i = blockIdx.z * kUpperThreads + tid; i < total; i += gridDim.z * kUpperThreads)
  {
    const int dy = bt + i / bw, dx = bl + i % bw;
    bool      in[5];
#pragma unroll
    for (int s = 0; s < 5; s++)
      in[s] = s_cu[s].idx >= 0 && dx >= s_cu[s].l && dx <= s_cu[s].r && dy >= s_cu[s].t && dy <= s_cu[s].b;
    uint32_t s64[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
      s64[j] = 0;
      if (in[j] || in[4])
      {
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
          const int  ri   = ((j >> 1) * 2 + (k >> 1)) * 4 + (j & 1) * 2 + (k & 1);
          const int4 info = s_reg[ri];
          if (s_surf[ri]) s64[j] += s_surf[ri][(size_t) (dy - info.y) * (info.z * 8) + (dx - info.x)];
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 5; s++)
    {
      if (!in[s]) continue;
      const uint32_t sad = s < 4 ? s64[s] : s64[0] + s64[1] + s64[2] + s64[3];
      if (sad <= key_cost(best[s]))
      {
        const uint32_t cost = sad + mv_cost(p.lambda, mv_bits_q(dx * 4, dy * 4, s_cu[s].pqx, s_cu[s].pqy, p.imvShift));
        const unsigned long long k = make_key(cost, dx, dy);
        if (k < best[s]) best[s] = k;
      }
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

size_t tree_sad_smem_bytes(int maxGx, int bandRows)
{
  return (size_t) kOffRef + (size_t) (bandRows + 31) * (size_t) (maxGx * 8 + 32) * 2;
}

cudaError_t launch_tree_sad(const TreeParams& p, int nPairs, cudaStream_t st)
{
  const size_t smem = tree_sad_smem_bytes(p.maxGx, p.bandRows);
  static size_t configured = 0;
  if (smem > configured)
  {
    cudaError_t e = cudaFuncSetAttribute(me_tree_sad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  dim3 grid(p.g.nRegX * p.g.nRegY, nPairs, 1);
  me_tree_sad_kernel<<<grid, kTreeThreads, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_tree_upper(const TreeParams& p, int nPairs, cudaStream_t st)
{
  dim3 grid(p.g.nCtuX * p.g.nCtuY, nPairs, 4);
  me_tree_upper_kernel<<<grid, kUpperThreads, 0, st>>>(p);
  return cudaGetLastError();
}

}   // namespace vtmme
