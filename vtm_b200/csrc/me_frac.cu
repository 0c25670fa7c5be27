// Fractional refinement + result write-out of the batched frame path.
//
// Same arithmetic as me_frac.cuh (xPatternSearchFracDIF, EncoderLib/InterSearch.cpp:4296-4338, with the separable
// 8-tap interpolation of InterpolationFilter.cpp:550-656 fused into the Hadamard SATD of RdCost.cpp:2267-2366),
// organised for throughput: the work item is one warp x one chunk of at most 16x16 samples, nothing but
// __syncwarp inside an item, four items per CTA.
//
//   8x8 and 16x16 CUs (94 % of the CUs) are one item each: half-pel stage, decision, quarter-pel stage, result.
//   Larger CUs are split into 16x16 chunks whose nine candidate SATDs are summed with atomics:
//     me_frac_items_kernel<0>  half-pel stage of every chunk            -> acc[cu][0..8]
//     me_frac_items_kernel<1>  decision from acc, quarter-pel stage      -> acc[cu][9..17]
//     me_frac_finish_kernel    decisions from acc, result
#include <cstdlib>

#include "me_frac.cuh"
#include "me_kernels.h"

#include "../../include/vtmme.h"

namespace vtmme {

namespace {

constexpr int kItemWarps  = 4;
constexpr int kChunk      = 16;
constexpr int kPatchW     = kChunk + 8;          // 24
constexpr int kPlaneStride = kChunk + 4;         // 20 words: rows of one 8-lane group fall into distinct 16-byte bank groups
constexpr int kPlaneElems = kPatchW * kPlaneStride;   // rows [-4, ch+4) x cols [0, cw)

struct __align__(16) ItemSmem
{
  uint32_t plane[3][kPlaneElems];   // 14-bit horizontal intermediates as VERTICAL pairs: word (r, c) = rows r (low half)
                                    // and r + 1 (high half) of column c — the operand layout of the 2-way dot product
  int16_t  org[kChunk * kChunk];
  uint16_t patch[kPatchW * kPatchW];
  uint32_t acc[12];
  uint32_t cfp[4][2];               // luma filter taps of the four quarter-pel phases, four signed bytes per word
};

struct ItemGeom
{
  int level, cu, x, y, size, chunkX, chunkY, nChunks;
};

// item index -> CU and chunk.  Items are ordered by level; level l >= 1 has (size/16)^2 chunks per CU.
__device__ __forceinline__ ItemGeom item_geom(const FrameGeom& g, int item)
{
  ItemGeom r;
  int base = 0;
#pragma unroll
  for (int l = 0; l < 5; l++)
  {
    const int per = l <= 1 ? 1 : (1 << (2 * (l - 1)));
    const int n   = g.nx[l] * g.ny[l] * per;
    if (item < base + n || l == 4)
    {
      const int li = (item - base) / per, ch = (item - base) - li * per;
      const int side = l <= 1 ? 1 : (1 << (l - 1));
      r.level   = l;
      r.cu      = g.off[l] + li;
      r.size    = 8 << l;
      r.x       = (li % g.nx[l]) * r.size;
      r.y       = (li / g.nx[l]) * r.size;
      r.chunkX  = (ch % side) * kChunk;
      r.chunkY  = (ch / side) * kChunk;
      r.nChunks = per;
      return r;
    }
    base += n;
  }
  return r;
}

__host__ __device__ inline int frame_items(const FrameGeom& g)
{
  int n = 0;
  for (int l = 0; l < 5; l++) n += g.nx[l] * g.ny[l] * (l <= 1 ? 1 : (1 << (2 * (l - 1))));
  return n;
}
__host__ __device__ inline int frame_items_small(const FrameGeom& g) { return g.nx[0] * g.ny[0] + g.nx[1] * g.ny[1]; }

// Nine candidate distortions of one CW x CW chunk (8x8 tiles) around quarter-pel offset (cqx,cqy) with step
// `step`, accumulated into sm.acc[0..8].  One warp.
// The reference narrows every filter output to Pel (int16).  With reference samples in [0, 2^bd) the narrowing never
// changes a value here, so it is not spelled out (one PRMT per sample saved): the first stage gives
// (sum - (8192 << shift)) >> shift with sum in [-24, 88] * (2^bd - 1), i.e. within +-14330 for bd = 8..10, and the second
// ((sum + offset) >> shift) stays within [-1055, 2079] before the clip.
template <bool HAD, int CW>
__device__ __forceinline__ void item_stage(ItemSmem& sm, const int16_t* __restrict__ org, int orgStride,
                                           const int16_t* __restrict__ refAtMv, int refStride, int cqx, int cqy, int step,
                                           int bitDepth, const int8_t (*tab)[2])
{
  constexpr int cw = CW, ch = CW, pw = CW + 8;
  const int lane = threadIdx.x & 31;
  const int hr   = max(2, 14 - bitDepth);
  const int maxv = (1 << bitDepth) - 1;
  if (lane < 12) sm.acc[lane] = 0;
  if (lane < 8)
  {
    const int16_t* cfs = c_lumaFilter[(lane >> 1) * 4] + (lane & 1) * 4;
    sm.cfp[lane >> 1][lane & 1] = (uint32_t) (cfs[0] & 0xff) | ((uint32_t) (cfs[1] & 0xff) << 8) | ((uint32_t) (cfs[2] & 0xff) << 16) |
                                  ((uint32_t) (cfs[3] & 0xff) << 24);
  }
  // 1. original chunk and reference patch rows/cols [-4, +4)
#pragma unroll
  for (int i = lane; i < cw * ch; i += 32)
  {
    const int y = i / cw, x = i % cw;
    sm.org[y * kChunk + x] = org[(size_t) y * orgStride + x];
  }
#pragma unroll 4
  for (int i = lane; i < (ch + 8) * pw; i += 32)
  {
    const int y = i / pw, x = i % pw;
    sm.patch[y * kPatchW + x] = (uint16_t) refAtMv[(ptrdiff_t) (y - 4) * refStride + (x - 4)];
  }
  __syncwarp();
  // 2. horizontal pass: lane r filters patch row r (picture row r-4) once for the three planes
  //    dqx = cqx + (p-1)*step, 14-bit intermediates (filter<8,false,true,false> / filterCopy<true,false>).  The eight
  //    taps of an output are four 2-way dot products (IDP.2A: two 16-bit samples x two 8-bit taps) on sample pairs —
  //    the even pairs are the row's own words, the odd pairs one funnel shift each.  A lane packs its row with the next
  //    lane's into vertical pairs for the second pass.
  {
    const int       rowl = min(lane, ch + 7);
    const uint32_t* prow = reinterpret_cast<const uint32_t*>(sm.patch + rowl * kPatchW);
    uint32_t        w[pw / 2], o[pw / 2];
#pragma unroll
    for (int j = 0; j < pw / 2; j++) w[j] = prow[j];
#pragma unroll
    for (int j = 0; j < pw / 2 - 1; j++) o[j] = __funnelshift_r(w[j], w[j + 1], 16);
    o[pw / 2 - 1] = 0;
#pragma unroll
    for (int p = 0; p < 3; p++)
    {
      const int dq = cqx + (p - 1) * step;
      const int ix = dq >> 2, px = dq & 3;   // ix in {-1, 0}
      int       v[cw];
      if (px == 0)
      {
#pragma unroll
        for (int c = 0; c < cw; c++)
        {
          const int s3 = (int) ((w[(c + 3) >> 1] >> (16 * ((c + 3) & 1))) & 0xffffu);
          const int s4 = (int) ((w[(c + 4) >> 1] >> (16 * ((c + 4) & 1))) & 0xffffu);
          v[c]         = ((ix ? s3 : s4) << hr) - 8192;
        }
      }
      else
      {
        const int cA = (int) sm.cfp[px][0], cB = (int) sm.cfp[px][1];
        const int shift = 6 - hr, off = 8192 << shift;
        if (ix)
        {
          // taps at samples c .. c+7: pairs starting at c (even c: own words, odd c: shifted words)
#pragma unroll
          for (int c = 0; c < cw; c++)
          {
            int sum = 0;
            if ((c & 1) == 0)
            {
              sum = __dp2a_lo((int) w[c / 2], cA, sum);
              sum = __dp2a_hi((int) w[c / 2 + 1], cA, sum);
              sum = __dp2a_lo((int) w[c / 2 + 2], cB, sum);
              sum = __dp2a_hi((int) w[c / 2 + 3], cB, sum);
            }
            else
            {
              sum = __dp2a_lo((int) o[c / 2], cA, sum);
              sum = __dp2a_hi((int) o[c / 2 + 1], cA, sum);
              sum = __dp2a_lo((int) o[c / 2 + 2], cB, sum);
              sum = __dp2a_hi((int) o[c / 2 + 3], cB, sum);
            }
            v[c] = (sum - off) >> shift;
          }
        }
        else
        {
          // taps at samples c+1 .. c+8
#pragma unroll
          for (int c = 0; c < cw; c++)
          {
            int sum = 0;
            if ((c & 1) == 1)
            {
              sum = __dp2a_lo((int) w[(c + 1) / 2], cA, sum);
              sum = __dp2a_hi((int) w[(c + 1) / 2 + 1], cA, sum);
              sum = __dp2a_lo((int) w[(c + 1) / 2 + 2], cB, sum);
              sum = __dp2a_hi((int) w[(c + 1) / 2 + 3], cB, sum);
            }
            else
            {
              sum = __dp2a_lo((int) o[c / 2], cA, sum);
              sum = __dp2a_hi((int) o[c / 2 + 1], cA, sum);
              sum = __dp2a_lo((int) o[c / 2 + 2], cB, sum);
              sum = __dp2a_hi((int) o[c / 2 + 3], cB, sum);
            }
            v[c] = (sum - off) >> shift;
          }
        }
      }
      uint32_t* dst = sm.plane[p] + rowl * kPlaneStride;
#pragma unroll
      for (int c = 0; c < cw; c += 4)
      {
        uint32_t q[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
          const int nxt = __shfl_down_sync(0xffffffffu, v[c + i], 1);
          q[i]          = ((uint32_t) v[c + i] & 0xffffu) | ((uint32_t) nxt << 16);
        }
        if (lane < ch + 7) *reinterpret_cast<uint4*>(dst + c) = make_uint4(q[0], q[1], q[2], q[3]);
      }
    }
  }
  __syncwarp();
  // 3. (candidate, 8x8 tile) units, 8 lanes each
  const int tilesX = cw >> 3, nTiles = tilesX * (ch >> 3), units = 9 * nTiles;
  const int group = lane >> 3, lit = lane & 7;
  for (int u0 = 0; u0 < units; u0 += 4)
  {
    const int  u      = u0 + group;
    const bool active = u < units;
    const int  uu     = active ? u : 0;
    const int  c = uu / nTiles, t = uu - c * nTiles;
    const int  tx = (t % tilesX) * 8, ty = (t / tilesX) * 8;
    const int  dqy = cqy + tab[c][1] * step;
    const int  iy = dqy >> 2, py = dqy & 3;
    const int  y = ty + lit;
    const uint32_t* pp = sm.plane[tab[c][0] + 1] + (y + iy + 4) * kPlaneStride + tx;   // pair row r: picture rows r-4, r-3
    int d[8];
    if (py == 0)
    {
      const uint4    a = *reinterpret_cast<const uint4*>(pp), b = *reinterpret_cast<const uint4*>(pp + 4);
      const uint32_t s[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
#pragma unroll
      for (int i = 0; i < 8; i++)
      {
        const int v = ((int) (short) (s[i] & 0xffffu) + 8192 + (1 << (hr - 1))) >> hr;
        d[i]        = min(max(v, 0), maxv);
      }
    }
    else
    {
      // taps 0..7 at picture rows y+iy-3 .. y+iy+4 = pair rows (y+iy+1), +2, +4, +6: four 2-way dot products per sample
      const int cA = (int) sm.cfp[py][0], cB = (int) sm.cfp[py][1];
      int       sum[8];
#pragma unroll
      for (int m = 0; m < 4; m++)
      {
        const uint32_t* row = pp + (2 * m - 3) * kPlaneStride;
        const uint4     a = *reinterpret_cast<const uint4*>(row), b = *reinterpret_cast<const uint4*>(row + 4);
        const uint32_t  s[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
          if (m == 0) sum[i] = __dp2a_lo((int) s[i], cA, 0);
          else if (m == 1) sum[i] = __dp2a_hi((int) s[i], cA, sum[i]);
          else if (m == 2) sum[i] = __dp2a_lo((int) s[i], cB, sum[i]);
          else sum[i] = __dp2a_hi((int) s[i], cB, sum[i]);
        }
      }
      const int shift = 6 + hr, offset = (1 << (shift - 1)) + (8192 << 6);
#pragma unroll
      for (int i = 0; i < 8; i++)
      {
        const int v = (sum[i] + offset) >> shift;
        d[i]        = min(max(v, 0), maxv);
      }
    }
    const uint4 ow = *reinterpret_cast<const uint4*>(sm.org + y * kChunk + tx);
    d[0] = (int) (short) (ow.x & 0xffffu) - d[0];
    d[1] = ((int) ow.x >> 16) - d[1];
    d[2] = (int) (short) (ow.y & 0xffffu) - d[2];
    d[3] = ((int) ow.y >> 16) - d[3];
    d[4] = (int) (short) (ow.z & 0xffffu) - d[4];
    d[5] = ((int) ow.z >> 16) - d[5];
    d[6] = (int) (short) (ow.w & 0xffffu) - d[6];
    d[7] = ((int) ow.w >> 16) - d[7];
    uint32_t v;
    if (HAD)
      v = satd_tile_rows<8, 8>(d, lit);
    else
    {
      v = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) v += (uint32_t) abs(d[i]);
#pragma unroll
      for (int m = 1; m < 8; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    }
    if (active && lit == 0) atomicAdd(&sm.acc[c], v);
  }
  __syncwarp();
}

__device__ __forceinline__ void run_stage(ItemSmem& sm, int useHad, int cw, const int16_t* org, int orgStride,
                                          const int16_t* rf, int refStride, int cqx, int cqy, int step, int bitDepth,
                                          const int8_t (*tab)[2])
{
  if (cw == 8)
  {
    if (useHad) item_stage<true, 8>(sm, org, orgStride, rf, refStride, cqx, cqy, step, bitDepth, tab);
    else item_stage<false, 8>(sm, org, orgStride, rf, refStride, cqx, cqy, step, bitDepth, tab);
  }
  else
  {
    if (useHad) item_stage<true, 16>(sm, org, orgStride, rf, refStride, cqx, cqy, step, bitDepth, tab);
    else item_stage<false, 16>(sm, org, orgStride, rf, refStride, cqx, cqy, step, bitDepth, tab);
  }
}

// xPatternRefinement's choice (InterSearch.cpp:727-756): first strict minimum over the nine candidates in list order.
// dist[] read by lanes 0..8 from `dist`; returns (direction, cost) to all lanes.
__device__ __forceinline__ void pick_best(const uint32_t* dist, const int8_t (*tab)[2], int baseQx, int baseQy, int step,
                                          int predQx, int predQy, double lambda, int& dir, uint32_t& cost)
{
  const int lane = threadIdx.x & 31;
  uint32_t  key  = 0xffffffffu;
  if (lane < 9)
  {
    const uint32_t c = dist[lane] + mv_cost(lambda, mv_bits_q(baseQx + tab[lane][0] * step, baseQy + tab[lane][1] * step,
                                                               predQx, predQy, 0));
    key = (c << 4) | (uint32_t) lane;   // costs stay below 2^28
  }
  key  = __reduce_min_sync(0xffffffffu, key);
  dir  = (int) (key & 15u);
  cost = key >> 4;
}

struct ItemParams
{
  FracFrameParams f;
  uint32_t*       acc;      // [nPairs][nCU][18] (only CUs of level >= 2 are used)
  int             nItems;   // per pair, for this launch
  int             itemBase; // first item index of this launch (0, or the first level-2 item)
};

template <int PASS>   // 0: every item, half stage (+ everything for single-item CUs); 1: quarter stage of multi-item CUs
__global__ void __launch_bounds__(kItemWarps * 32) me_frac_items_kernel(ItemParams ip)
{
  __shared__ ItemSmem s_items[kItemWarps];
  const FracFrameParams& p = ip.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kItemWarps + warp;
  if (item >= ip.nItems) return;
  const int      pair = blockIdx.y;
  const int      nCU  = p.g.off[5];
  ItemSmem&      sm   = s_items[warp];
  const ItemGeom ig   = item_geom(p.g, ip.itemBase + item);

  const unsigned long long key = p.keys[(size_t) pair * nCU + ig.cu];
  short2 pr = make_short2(0, 0);
  if (p.predQ) pr = p.predQ[(size_t) pair * nCU + ig.cu];
  const int    dx = key_dx(key), dy = key_dy(key);
  const DevPic cur = p.cur[pair], ref = p.ref[pair];
  const int    cw = min(ig.size, kChunk);
  const int16_t* org = cur.origin + (size_t) (ig.y + ig.chunkY) * cur.stride + ig.x + ig.chunkX;
  const int16_t* rf  = ref.origin + (ptrdiff_t) (ig.y + ig.chunkY + dy) * ref.stride + (ig.x + ig.chunkX + dx);
  uint32_t*      acc = ip.acc + ((size_t) pair * nCU + ig.cu) * 18;

  int      hdir = 0, qdir = 0;
  uint32_t cost = 0;
  if (PASS == 0)
  {
    run_stage(sm, p.useHad, cw, org, cur.stride, rf, ref.stride, 0, 0, 2, p.bitDepth, c_refineH);
    if (ig.nChunks > 1)
    {
      if (lane < 9) atomicAdd(&acc[lane], sm.acc[lane]);
      return;
    }
    pick_best(sm.acc, c_refineH, dx * 4, dy * 4, 2, pr.x, pr.y, p.lambda, hdir, cost);
  }
  else
  {
    pick_best(acc, c_refineH, dx * 4, dy * 4, 2, pr.x, pr.y, p.lambda, hdir, cost);
  }
  const int hx = c_refineH[hdir][0], hy = c_refineH[hdir][1];
  if (p.imvShift == 0)
  {
    __syncwarp();
    run_stage(sm, p.useHad, cw, org, cur.stride, rf, ref.stride, hx * 2, hy * 2, 1, p.bitDepth, c_refineQ);
    if (ig.nChunks > 1)
    {
      if (lane < 9) atomicAdd(&acc[9 + lane], sm.acc[lane]);
      return;
    }
    pick_best(sm.acc, c_refineQ, dx * 4 + hx * 2, dy * 4 + hy * 2, 1, pr.x, pr.y, p.lambda, qdir, cost);
  }
  else if (ig.nChunks > 1)
    return;
  if (lane == 0)
  {
    const int qx = p.imvShift == 0 ? c_refineQ[qdir][0] : 0, qy = p.imvShift == 0 ? c_refineQ[qdir][1] : 0;
    vtmme_cu_result res;
    res.intX     = (int16_t) dx;
    res.intY     = (int16_t) dy;
    res.intSad   = key_cost(key) - mv_cost(p.lambda, mv_bits_q(dx * 4, dy * 4, pr.x, pr.y, p.imvShift));
    res.mvQx     = (int16_t) (dx * 4 + hx * 2 + qx);
    res.mvQy     = (int16_t) (dy * 4 + hy * 2 + qy);
    res.fracCost = cost;
    reinterpret_cast<vtmme_cu_result*>(p.results)[(size_t) pair * nCU + ig.cu] = res;
  }
}

// Results of the CUs whose stages were accumulated across chunks (levels >= 2), and of every CU when fracMode == 0.
__global__ void __launch_bounds__(128) me_frac_finish_kernel(ItemParams ip, int firstCu)
{
  const FracFrameParams& p = ip.f;
  const int nCU = p.g.off[5];
  const int cu  = firstCu + blockIdx.x * 4 + (threadIdx.x >> 5);
  if (cu >= nCU) return;
  const int pair = blockIdx.y, lane = threadIdx.x & 31;
  const unsigned long long key = p.keys[(size_t) pair * nCU + cu];
  short2 pr = make_short2(0, 0);
  if (p.predQ) pr = p.predQ[(size_t) pair * nCU + cu];
  const int      dx = key_dx(key), dy = key_dy(key);
  const uint32_t intCost = key_cost(key);
  int      hdir = 0, qdir = 0;
  uint32_t cost = intCost;
  int      hx = 0, hy = 0, qx = 0, qy = 0;
  if (p.fracMode)
  {
    const uint32_t* acc = ip.acc + ((size_t) pair * nCU + cu) * 18;
    pick_best(acc, c_refineH, dx * 4, dy * 4, 2, pr.x, pr.y, p.lambda, hdir, cost);
    hx = c_refineH[hdir][0];
    hy = c_refineH[hdir][1];
    if (p.imvShift == 0)
    {
      pick_best(acc + 9, c_refineQ, dx * 4 + hx * 2, dy * 4 + hy * 2, 1, pr.x, pr.y, p.lambda, qdir, cost);
      qx = c_refineQ[qdir][0];
      qy = c_refineQ[qdir][1];
    }
  }
  if (lane == 0)
  {
    vtmme_cu_result res;
    res.intX     = (int16_t) dx;
    res.intY     = (int16_t) dy;
    res.intSad   = intCost - mv_cost(p.lambda, mv_bits_q(dx * 4, dy * 4, pr.x, pr.y, p.imvShift));
    res.mvQx     = (int16_t) (dx * 4 + hx * 2 + qx);
    res.mvQy     = (int16_t) (dy * 4 + hy * 2 + qy);
    res.fracCost = cost;
    reinterpret_cast<vtmme_cu_result*>(p.results)[(size_t) pair * nCU + cu] = res;
  }
}

}   // namespace

size_t frac_frame_acc_bytes(const FrameGeom& g, int nPairs) { return (size_t) nPairs * g.off[5] * 18 * sizeof(uint32_t); }

cudaError_t launch_frac_frame(const FracFrameParams& p, uint32_t* acc, int nPairs, cudaStream_t st, int* launches)
{
  ItemParams ip;
  ip.f   = p;
  ip.acc = acc;
  cudaError_t e;
  if (!p.fracMode)
  {
    ip.nItems = ip.itemBase = 0;
    dim3 g((p.g.off[5] + 3) / 4, nPairs, 1);
    me_frac_finish_kernel<<<g, 128, 0, st>>>(ip, 0);
    *launches += 1;
    return cudaGetLastError();
  }
  // default: one thread per 8x8 tile (me_frac_tile.cu); VTMME_FRAC_VARIANT=items keeps the warp-per-chunk kernels below
  static const bool useItems = [] { const char* v = getenv("VTMME_FRAC_VARIANT"); return v && v[0] == 'i'; }();
  if (!useItems) return launch_frac_frame_tiles(p, nPairs, st, launches);
  if ((e = cudaMemsetAsync(acc, 0, frac_frame_acc_bytes(p.g, nPairs), st)) != cudaSuccess) return e;
  const int nAll = frame_items(p.g), nSmall = frame_items_small(p.g);
  ip.nItems   = nAll;
  ip.itemBase = 0;
  dim3 g0((nAll + kItemWarps - 1) / kItemWarps, nPairs, 1);
  me_frac_items_kernel<0><<<g0, kItemWarps * 32, 0, st>>>(ip);
  *launches += 1;
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (nAll > nSmall)
  {
    if (p.imvShift == 0)
    {
      ip.nItems   = nAll - nSmall;
      ip.itemBase = nSmall;
      dim3 g1((ip.nItems + kItemWarps - 1) / kItemWarps, nPairs, 1);
      me_frac_items_kernel<1><<<g1, kItemWarps * 32, 0, st>>>(ip);
      *launches += 1;
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    const int firstCu = p.g.off[2];
    dim3 g2((p.g.off[5] - firstCu + 3) / 4, nPairs, 1);
    me_frac_finish_kernel<<<g2, 128, 0, st>>>(ip, firstCu);
    *launches += 1;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}   // namespace vtmme
