// Fractional refinement of the batched frame path, one THREAD per 8x8 SATD tile.
//
// Arithmetic: xPatternSearchFracDIF (EncoderLib/InterSearch.cpp:4296-4338) in the direct form of me_frac.cuh — the
// candidate at quarter-pel offset (dqx, dqy) from the integer MV is the reference interpolated at integer base dq >> 2
// with phase dq & 3, horizontal 8-tap pass into 14-bit intermediates (InterpolationFilter::filter<8,false,true,false>,
// CommonLib/InterpolationFilter.cpp:550-656), vertical pass with the final rounding and clip (filter<8,true,false,true>),
// Hadamard SATD of the 8x8 tiles (RdCost::xCalcHADs8x8, CommonLib/RdCost.cpp:2267-2366) — nine candidates per stage in the
// order of s_acMvRefineH / s_acMvRefineQ, first strict minimum (xPatternRefinement, InterSearch.cpp:707-761).  A phase of 0
// is the same filter with taps {0,0,0,64,0,0,0,0}: (64 s - (8192 << sh)) >> sh == (s << hr) - 8192 and
// (64 h + 2^(5+hr) + (8192 << 6)) >> (6+hr) == (h + 8192 + 2^(hr-1)) >> hr, i.e. filterCopy's results, so no candidate
// takes a different code path.  The centre of the quarter-pel stage is the winner of the half-pel stage; its distortion
// is taken from there instead of being computed again.
//
// Organisation (why it is faster than one warp per 16x16 chunk, me_frac.cu):
//   * a tile never leaves its thread: the 2-D Hadamard is 6 x 64 register butterflies, no shuffles;
//   * two pipes: the filters are 2-way dot products (IDP.2A: two 16-bit samples x two 8-bit taps) on the integer pipe, the
//     difference, the butterflies and the sum of magnitudes are FADDs on the FP32 pipe.  Samples and coefficients below 2^24
//     are *denormal / small normal* floats whose bit pattern is the integer (0x00800000 + m is (2^23 + m) * 2^-149), FADD on
//     them is exact fixed-point arithmetic at full rate (the tree kernel's trick), negation and magnitude are operand
//     modifiers;
//   * the horizontally filtered rows of a tile go through a thread-private strip of shared memory (16 rows x 8 samples),
//     never through another thread, so a CU needs block-level synchronisation only for the sums of its tiles.
// The reference patch of a CU ((S+8)^2 samples at the integer MV) is staged once per CU by the whole CTA.
#include "me_frac.cuh"
#include "me_kernels.h"

#include "../../include/vtmme.h"

namespace vtmme {

namespace {

__host__ __device__ constexpr uint32_t pack4(int a, int b, int c, int d)
{
  return (uint32_t) (a & 0xff) | ((uint32_t) (b & 0xff) << 8) | ((uint32_t) (c & 0xff) << 16) | ((uint32_t) (d & 0xff) << 24);
}
// luma taps of the quarter-pel phases 0, 1/4, 1/2, 3/4 (c_lumaFilter rows 0, 4, 8, 12; InterpolationFilter.cpp:77-95),
// four signed bytes per word: [phase][0] = taps 0..3, [phase][1] = taps 4..7
__constant__ uint32_t c_tapW[4][2] = { { pack4(0, 0, 0, 64), pack4(0, 0, 0, 0) },
                                           { pack4(-1, 4, -10, 58), pack4(17, -5, 1, 0) },
                                           { pack4(-1, 4, -11, 40), pack4(40, -11, 4, -1) },
                                           { pack4(0, 1, -5, 17), pack4(58, -10, 4, -1) } };
// (dx + 1) * 3 + (dy + 1) -> position in s_acMvRefineH / s_acMvRefineQ (me_frac.cuh c_refineH / c_refineQ)
__constant__ int8_t c_orderH[9] = { 5, 3, 7, 1, 0, 2, 6, 4, 8 };
__constant__ int8_t c_orderQ[9] = { 3, 5, 7, 1, 0, 2, 4, 6, 8 };

template <int LEVEL>
struct TileCfg
{
  static constexpr int S        = 8 << LEVEL;                 // CU size
  static constexpr int TPR      = S / 8;                      // tiles per CU row
  static constexpr int TPC      = TPR * TPR;                  // tiles (threads) per CU
  static constexpr int THREADS  = LEVEL == 4 ? 256 : 128;
  static constexpr int CPC      = THREADS / TPC;              // CUs per CTA
  static constexpr int ROWS     = S + 8;                      // patch rows
  static constexpr int PWW      = (S + 8) / 2;                // patch row stride in words (a multiple of 4)
  // CUs whose tiles share a quarter-warp are staggered by 16 bytes x tiles per row: the eight 128-bit reads of a
  // quarter-warp fall into distinct bank groups
  static constexpr int CU_WORDS = ROWS * PWW + (LEVEL < 3 ? 4 * TPR : 0);
  static constexpr int PLANE_Q  = 17;                         // private strip: 16 rows of 16 bytes + 16 bytes of stagger
  static constexpr size_t SMEM  = (size_t) CPC * CU_WORDS * 4 + (size_t) THREADS * PLANE_Q * 16 + (size_t) CPC * (8 + (LEVEL >= 3 ? 18 * 4 : 0));
};

// thread -> (CU within the CTA, tile column, tile row); see CU_WORDS for the bank argument
template <int LEVEL>
__device__ __forceinline__ void tile_of_thread(int t, int& cu, int& tx, int& ty)
{
  constexpr int TXB = LEVEL < 3 ? LEVEL : 3;
  const int lane8 = t & 7, rest = t >> 3;
  tx = lane8 & ((1 << TXB) - 1);
  const int cuLo = lane8 >> TXB;
  if (LEVEL == 4)
  {
    tx |= (rest & 1) << 3;
    ty = rest >> 1;
    cu = 0;
  }
  else
  {
    ty = rest & ((1 << LEVEL) - 1);
    cu = ((rest >> LEVEL) << (3 - TXB)) | cuLo;
  }
}

// sum over the tiles of a CU that share the warp (lane bits of tx and ty, see tile_of_thread)
template <int LEVEL>
__device__ __forceinline__ uint32_t warp_sum_cu(uint32_t v)
{
  if (LEVEL == 1)
  {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
  }
  else if (LEVEL == 2)
  {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
  }
  else if (LEVEL >= 3)
  {
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  }
  return v;
}

// minimum over the 4 / 16 lanes that share the decision of a CU
template <int LEVEL>
__device__ __forceinline__ uint32_t warp_min_cu(uint32_t v)
{
  if (LEVEL == 1)
  {
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 8));
  }
  else if (LEVEL == 2)
  {
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 8));
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 16));
  }
  else if (LEVEL >= 3)
  {
#pragma unroll
    for (int m = 1; m < 16; m <<= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, m));
  }
  return v;
}

// Horizontal pass of one plane (candidate column offset dqx = 4 ix + px) for the 16 rows x 8 columns a tile needs:
// patch rows ty*8 .. ty*8+15 (picture rows -4 .. +11 of the tile), output column c = taps at patch columns c+ix+1 .. c+ix+8.
__device__ __forceinline__ void tile_hpass(const uint32_t* __restrict__ patchTile, int rowWords, uint4* __restrict__ strip,
                                           int ix, int px, int hr)
{
  const int      cA = (int) c_tapW[px][0], cB = (int) c_tapW[px][1];
  const int      shift = 6 - hr, negOff = -(8192 << shift);
  const uint32_t sh = (uint32_t) (ix + 1) * 16u;   // window = patch columns ix+1 ..: 0 or one sample to the right
#pragma unroll 4   // not 16: the kernel's code should stay well inside the instruction cache
  for (int r = 0; r < 16; r++)
  {
    const uint4    a = *reinterpret_cast<const uint4*>(patchTile + r * rowWords);
    const uint4    b = *reinterpret_cast<const uint4*>(patchTile + r * rowWords + 4);
    const uint32_t W[9] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, 0u };
    uint32_t       A[8], B[7];
#pragma unroll
    for (int j = 0; j < 8; j++) A[j] = __funnelshift_r(W[j], W[j + 1], sh);       // window columns 2j, 2j+1
#pragma unroll
    for (int j = 0; j < 7; j++) B[j] = __funnelshift_r(A[j], A[j + 1], 16);       // window columns 2j+1, 2j+2
    int v[8];
#pragma unroll
    for (int c = 0; c < 8; c++)
    {
      const uint32_t* s = (c & 1) ? &B[(c - 1) / 2] : &A[c / 2];
      int sum = __dp2a_lo((int) s[0], cA, negOff);
      sum     = __dp2a_hi((int) s[1], cA, sum);
      sum     = __dp2a_lo((int) s[2], cB, sum);
      sum     = __dp2a_hi((int) s[3], cB, sum);
      v[c]    = sum >> shift;
    }
    strip[r] = make_uint4(__byte_perm(v[0], v[1], 0x5410), __byte_perm(v[2], v[3], 0x5410), __byte_perm(v[4], v[5], 0x5410),
                          __byte_perm(v[6], v[7], 0x5410));
  }
}

// Vertical pass + clip of one candidate (row offset dqy = 4 iy + py) out of the thread's strip, difference against the
// original tile and distortion (HAD: 8x8 Hadamard SATD; else SAD), all 64 samples in registers.
template <bool HAD>
__device__ __forceinline__ uint32_t tile_candidate(const uint4* __restrict__ strip, const float (&org)[64], int iy, int py, int hr,
                                                   int maxv)
{
  const int cA = (int) c_tapW[py][0], cB = (int) c_tapW[py][1];
  const int shift = 6 + hr, offset = (1 << (shift - 1)) + (8192 << 6);
  const uint4* rows = strip + (iy + 1);   // rows[q]: picture row (q + iy - 3) of the tile; output y uses rows[y .. y+7]
  int   sum[64];
  uint4 prev = rows[0];
#pragma unroll
  for (int q = 0; q < 14; q++)
  {
    const uint4    cur = rows[q + 1];
    const uint32_t pw[4] = { prev.x, prev.y, prev.z, prev.w }, cw[4] = { cur.x, cur.y, cur.z, cur.w };
    uint32_t       P[8];   // column c: rows q (low half) and q + 1 (high half)
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
      P[2 * j]     = __byte_perm(pw[j], cw[j], 0x5410);
      P[2 * j + 1] = __byte_perm(pw[j], cw[j], 0x7632);
    }
#pragma unroll
    for (int m = 0; m < 4; m++)   // tap pair m of output row y = q - 2m
    {
      const int y = q - 2 * m;
      if (y < 0 || y > 7) continue;
#pragma unroll
      for (int c = 0; c < 8; c++)
      {
        int& s = sum[y * 8 + c];
        if (m == 0) s = __dp2a_lo((int) P[c], cA, offset);
        else if (m == 1) s = __dp2a_hi((int) P[c], cA, s);
        else if (m == 2) s = __dp2a_lo((int) P[c], cB, s);
        else s = __dp2a_hi((int) P[c], cB, s);
      }
    }
    prev = cur;
  }
  float f[64];
#pragma unroll
  for (int i = 0; i < 64; i++)
  {
    const int v = __vimin_s32_relu(sum[i] >> shift, maxv);   // clip to [0, maxv] in one instruction
    f[i]        = org[i] - __int_as_float(v);   // exact: both are integers below 2^24 carried as (de)normal floats
  }
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (HAD)
  {
#pragma unroll
    for (int len = 1; len < 64; len <<= 1)
#pragma unroll
      for (int i = 0; i < 64; i += len << 1)
#pragma unroll
        for (int k = i; k < i + len; k++)
        {
          const float a = f[k], b = f[k + len];
          f[k]       = a + b;
          f[k + len] = a - b;
        }
#pragma unroll
    for (int i = 4; i < 64; i += 4)
    {
      s0 += fabsf(f[i]);
      s1 += fabsf(f[i + 1]);
      s2 += fabsf(f[i + 2]);
      s3 += fabsf(f[i + 3]);
    }
    s1 += fabsf(f[1]);
    s2 += fabsf(f[2]);
    s3 += fabsf(f[3]);
    const uint32_t dc = __float_as_uint(fabsf(f[0]));
    const uint32_t s  = __float_as_uint((s0 + s1) + (s2 + s3)) + (dc >> 2);   // DC term scaled (JVET_R0164, RdCost.cpp:2354-2357)
    return (s + 2) >> 2;
  }
#pragma unroll
  for (int i = 0; i < 64; i += 4)
  {
    s0 += fabsf(f[i]);
    s1 += fabsf(f[i + 1]);
    s2 += fabsf(f[i + 2]);
    s3 += fabsf(f[i + 3]);
  }
  return __float_as_uint((s0 + s1) + (s2 + s3));
}

template <int LEVEL, bool HAD>
__global__ void __launch_bounds__(TileCfg<LEVEL>::THREADS) me_frac_tile_kernel(FracFrameParams p)
{
  using Cfg = TileCfg<LEVEL>;
  extern __shared__ __align__(16) unsigned char s_raw[];
  uint32_t* sPatch = reinterpret_cast<uint32_t*>(s_raw);
  uint4*    sStrip = reinterpret_cast<uint4*>(s_raw + (size_t) Cfg::CPC * Cfg::CU_WORDS * 4);
  int2*     sMv    = reinterpret_cast<int2*>(s_raw + (size_t) Cfg::CPC * Cfg::CU_WORDS * 4 + (size_t) Cfg::THREADS * Cfg::PLANE_Q * 16);
  uint32_t* sAcc   = reinterpret_cast<uint32_t*>(sMv + Cfg::CPC);   // [CPC][18], levels 3 and 4 only

  const int t = threadIdx.x, pair = blockIdx.y;
  const int nCU = p.g.off[5], nLevel = p.g.nx[LEVEL] * p.g.ny[LEVEL];
  const int cuBase = blockIdx.x * Cfg::CPC;
  const DevPic cur = p.cur[pair], ref = p.ref[pair];

  // integer MVs of the CTA's CUs
  if (t < Cfg::CPC)
  {
    int2 mv = make_int2(0, 0);
    if (cuBase + t < nLevel)
    {
      const unsigned long long key = p.keys[(size_t) pair * nCU + p.g.off[LEVEL] + cuBase + t];
      if (key != ~0ull) mv = make_int2(key_dx(key), key_dy(key));   // ~0: no candidate (the search reported an error)
    }
    sMv[t] = mv;
  }
  if (LEVEL >= 3)
    for (int i = t; i < Cfg::CPC * 18; i += Cfg::THREADS) sAcc[i] = 0;
  __syncthreads();
  int cu, tx, ty;
  tile_of_thread<LEVEL>(t, cu, tx, ty);
  const int  li    = cuBase + cu;
  const bool valid = li < nLevel;
  const int  liC   = valid ? li : 0;
  const int  x = (liC % p.g.nx[LEVEL]) * Cfg::S, y = (liC / p.g.nx[LEVEL]) * Cfg::S;
  const int  cuIdx = p.g.off[LEVEL] + liC;
  short2 pr = make_short2(0, 0);
  if (p.predQ) pr = p.predQ[(size_t) pair * nCU + cuIdx];
  // the thread's original tile as integer-valued floats (samples of a picture are >= 0); the loads are issued before the
  // patch is staged so that their latency is hidden behind it
  uint4 orgRaw[8];
  {
    const int16_t* o = cur.origin + (size_t) (y + ty * 8) * cur.stride + x + tx * 8;
#pragma unroll
    for (int r = 0; r < 8; r++) orgRaw[r] = *reinterpret_cast<const uint4*>(o + (size_t) r * cur.stride);
  }
  // reference patches: rows -4 .. S+3, columns -4 .. S+3 around the block at the integer MV, 16-bit pairs.  One unit = 16
  // bytes of a patch row, assembled from five aligned words (a patch may start at an odd sample); several units in flight
  // per thread.
  {
    constexpr int UPR = Cfg::PWW / 4, UNITS = Cfg::CPC * Cfg::ROWS * UPR;
    const int lastCu = nLevel - 1 - cuBase;   // CUs past the end of the level re-read the last one (never used)
#pragma unroll 4
    for (int i = t; i < UNITS; i += Cfg::THREADS)
    {
      const int c = i / (Cfg::ROWS * UPR), rem = i - c * (Cfg::ROWS * UPR);
      const int row = rem / UPR, u = rem - row * UPR;
      const int cc = min(c, lastCu), lj = cuBase + cc;
      const int2 cmv = sMv[cc];
      const int  px = (lj % p.g.nx[LEVEL]) * Cfg::S, py = (lj / p.g.nx[LEVEL]) * Cfg::S;
      const int16_t*  g  = ref.origin + (ptrdiff_t) (py + cmv.y - 4 + row) * ref.stride + (px + cmv.x - 4) + 8 * u;
      const uint32_t* gw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(g) & ~(uintptr_t) 3);
      const uint32_t  sh = (uint32_t) (reinterpret_cast<uintptr_t>(g) & 2) * 8u;
      const uint32_t  w0 = gw[0], w1 = gw[1], w2 = gw[2], w3 = gw[3], w4 = gw[4];
      *reinterpret_cast<uint4*>(sPatch + c * Cfg::CU_WORDS + row * Cfg::PWW + 4 * u) =
        make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
    }
  }
  __syncthreads();
  const int2 mv = sMv[cu];
  float org[64];
#pragma unroll
  for (int r = 0; r < 8; r++)
  {
    const uint32_t w[4] = { orgRaw[r].x, orgRaw[r].y, orgRaw[r].z, orgRaw[r].w };
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
      org[r * 8 + 2 * j]     = __uint_as_float(w[j] & 0xffffu);
      org[r * 8 + 2 * j + 1] = __uint_as_float(w[j] >> 16);
    }
  }
  const uint32_t* patchTile = sPatch + cu * Cfg::CU_WORDS + (ty * 8) * Cfg::PWW + tx * 4;
  uint4*          strip     = sStrip + (size_t) t * Cfg::PLANE_Q;
  const int hr = max(2, 14 - p.bitDepth), maxv = (1 << p.bitDepth) - 1;
  const int tileInCu = ty * Cfg::TPR + tx;
  constexpr int G = Cfg::TPC < 16 ? Cfg::TPC : 16;   // lanes sharing the decision of a CU

  int      cqx = 0, cqy = 0;          // centre of the stage, quarter-pel units relative to the integer MV
  uint32_t bestCost = 0, centreDist = 0;
  const int nStages = p.imvShift == 0 ? 2 : 1;
#pragma unroll 1
  for (int stage = 0; stage < nStages; stage++)
  {
    const int     step  = stage == 0 ? 2 : 1;
    const int8_t* order = stage == 0 ? c_orderH : c_orderQ;
    uint32_t      dist[9];
#pragma unroll
    for (int c = 0; c < 9; c++) dist[c] = 0;
#pragma unroll 1
    for (int pl = 0; pl < 3; pl++)
    {
      const int dqx = cqx + (pl - 1) * step;
      tile_hpass(patchTile, Cfg::PWW, strip, dqx >> 2, dqx & 3, hr);
#pragma unroll 1
      for (int v = 0; v < 3; v++)
      {
        if (stage == 1 && pl == 1 && v == 1) continue;   // the centre: distortion known from the half-pel stage
        const int      dqy = cqy + (v - 1) * step;
        const uint32_t d   = tile_candidate<HAD>(strip, org, dqy >> 2, dqy & 3, hr, maxv);
        const int      idx = order[pl * 3 + v];
#pragma unroll
        for (int c = 0; c < 9; c++) dist[c] = c == idx ? d : dist[c];
      }
    }
    // sums over the tiles of the CU
    if (LEVEL >= 1)
    {
#pragma unroll
      for (int c = 0; c < 9; c++) dist[c] = warp_sum_cu<LEVEL>(dist[c]);
    }
    if (LEVEL >= 3)
    {
      if ((t & 31) == 0)
#pragma unroll
        for (int c = 0; c < 9; c++) atomicAdd(&sAcc[cu * 18 + stage * 9 + c], dist[c]);
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 9; c++) dist[c] = sAcc[cu * 18 + stage * 9 + c];
    }
    if (stage == 1) dist[0] = centreDist;
    // xPatternRefinement's choice: first strict minimum in list order; the candidates are spread over G lanes of the CU
    const int8_t (*tab)[2] = stage == 0 ? c_refineH : c_refineQ;
    uint32_t key = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < 9; c++)
    {
      if ((c & (G - 1)) != (tileInCu & (G - 1))) continue;
      const uint32_t cost = dist[c] + mv_cost(p.lambda, mv_bits_q(mv.x * 4 + cqx + tab[c][0] * step, mv.y * 4 + cqy + tab[c][1] * step,
                                                                  pr.x, pr.y, 0));
      key = min(key, (cost << 4) | (uint32_t) c);   // costs stay below 2^28
    }
    key = warp_min_cu<LEVEL>(key);
    const int dir = (int) (key & 15u);
    bestCost      = key >> 4;
    centreDist    = 0;
#pragma unroll
    for (int c = 0; c < 9; c++) centreDist = c == dir ? dist[c] : centreDist;
    cqx += tab[dir][0] * step;
    cqy += tab[dir][1] * step;
  }
  if (valid && tileInCu == 0)
  {
    const unsigned long long key = p.keys[(size_t) pair * nCU + cuIdx];
    vtmme_cu_result res;
    res.intX     = (int16_t) mv.x;
    res.intY     = (int16_t) mv.y;
    res.intSad   = key_cost(key) - mv_cost(p.lambda, mv_bits_q(mv.x * 4, mv.y * 4, pr.x, pr.y, p.imvShift));
    res.mvQx     = (int16_t) (mv.x * 4 + cqx);
    res.mvQy     = (int16_t) (mv.y * 4 + cqy);
    res.fracCost = bestCost;
    reinterpret_cast<vtmme_cu_result*>(p.results)[(size_t) pair * nCU + cuIdx] = res;
  }
}

template <int LEVEL, bool HAD>
cudaError_t launch_level(const FracFrameParams& p, int nPairs, cudaStream_t st, int* launches)
{
  using Cfg = TileCfg<LEVEL>;
  static SmemOptIn optIn;
  const int n = p.g.nx[LEVEL] * p.g.ny[LEVEL];
  if (n == 0) return cudaSuccess;
  cudaError_t e = optIn.ensure(me_frac_tile_kernel<LEVEL, HAD>, Cfg::SMEM);
  if (e != cudaSuccess) return e;
  dim3 grid((n + Cfg::CPC - 1) / Cfg::CPC, nPairs, 1);
  me_frac_tile_kernel<LEVEL, HAD><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
  *launches += 1;
  return cudaGetLastError();
}

template <bool HAD>
cudaError_t launch_all(const FracFrameParams& p, int nPairs, cudaStream_t st, int* launches)
{
  cudaError_t e;
  if ((e = launch_level<4, HAD>(p, nPairs, st, launches)) != cudaSuccess) return e;
  if ((e = launch_level<3, HAD>(p, nPairs, st, launches)) != cudaSuccess) return e;
  if ((e = launch_level<2, HAD>(p, nPairs, st, launches)) != cudaSuccess) return e;
  if ((e = launch_level<1, HAD>(p, nPairs, st, launches)) != cudaSuccess) return e;
  return launch_level<0, HAD>(p, nPairs, st, launches);
}

}   // namespace

// fracMode != 0: one launch per CU level, every CU refined (both stages) and its result written
cudaError_t launch_frac_frame_tiles(const FracFrameParams& p, int nPairs, cudaStream_t st, int* launches)
{
  return p.useHad ? launch_all<true>(p, nPairs, st, launches) : launch_all<false>(p, nPairs, st, launches);
}

}   // namespace vtmme
