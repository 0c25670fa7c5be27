// TZ search, InterSearch::xTZSearch (EncoderLib/InterSearch.cpp:3640-3974) with xTZSearchHelp (:330-417),
// xTZ2PointSearch (:420-446) and xTZ8PointDiamondSearch (:503-705): the integer search of FastSearch=1
// (MESEARCH_DIAMOND), FastSearch=3 (MESEARCH_DIAMOND_ENHANCED, `extended`) and of the cached-MV re-search (`fast`).
//
// The reference probes one position after the other and updates its best state after each.  Between two decisions
// that depend on that state, however, the probe positions are fixed: the points of one diamond, of the two-point
// step, of the raster scan, the history seeds.  Such a *batch* is evaluated here in parallel — one warp per point, the
// lanes over the samples — and reduced with the key (cost, order in the batch): the first strict minimum in
// probe order, which is exactly the state the reference's sequential updates end in.  WARPS warps run one search
// (4 for a per-call job, where latency counts; 1 in the batched frame search, where a CTA carries four independent
// searches and nothing but warp-level synchronisation is needed); all their threads execute the (uniform) control
// flow, the state lives in registers.
#pragma once
#include "me_common.cuh"
#include "me_kernels.h"

namespace vtmme {

constexpr int kTzThreads   = 128;   // per-call jobs: 4 warps per search
constexpr int kTzMaxPoints = 16;    // a diamond
constexpr int kTzChunk     = 256;   // raster points per reduction

struct TzState
{
  uint32_t best;          // cStruct.uiBestSad (cost incl. rate); 0xffffffff = none yet
  int      bx, by;        // iBestX, iBestY
  uint32_t dist, round;   // uiBestDistance, uiBestRound
  int      pnr;           // ucPointNr
  int      l, r, t, b;    // cStruct.searchRange
};

struct TzSmem
{
  unsigned long long key;
  uint32_t           gate[kTzChunk], full[kTzChunk];   // staged probes of a batch: acceptance threshold, exact cost
};

// SAD of the pattern (row stride patStride, 8-byte aligned rows) against the reference block at `cur`, rows r with
// (r & (step-1)) == 0, shifted back up: DistParam::subShift semantics (RdCost.cpp:493-528).  One warp; the sum is
// returned in every lane.
__device__ __forceinline__ uint32_t tz_warp_sad(const int16_t* pat, int patStride, int w, int h, const int16_t* cur,
                                                int refStride, int subShift)
{
  const int lane = threadIdx.x & 31;
  const int q = w >> 2, rows = h >> subShift, units = q * rows;
  const bool even = (reinterpret_cast<uintptr_t>(cur) & 2) == 0;   // reference rows start on a 4-byte boundary (stride is even)
  uint32_t  s = 0;
  for (int u = lane; u < units; u += 32)
  {
    const int      r = (u / q) << subShift, c = (u % q) << 2;
    const uint2    o = *reinterpret_cast<const uint2*>(pat + r * patStride + c);
    const int16_t* p = cur + (ptrdiff_t) r * refStride + c;
    int            p0, p1, p2, p3;
    if (even)
    {
      const uint32_t a = *reinterpret_cast<const uint32_t*>(p), b = *reinterpret_cast<const uint32_t*>(p + 2);
      p0 = (int) (short) (a & 0xffffu);
      p1 = (int) a >> 16;
      p2 = (int) (short) (b & 0xffffu);
      p3 = (int) b >> 16;
    }
    else
    {
      const uint32_t m = *reinterpret_cast<const uint32_t*>(p + 1);
      p0 = (int) p[0];
      p1 = (int) (short) (m & 0xffffu);
      p2 = (int) m >> 16;
      p3 = (int) p[3];
    }
    s = __sad((int) (short) (o.x & 0xffffu), p0, s);
    s = __sad((int) o.x >> 16, p1, s);
    s = __sad((int) (short) (o.y & 0xffffu), p2, s);
    s = __sad((int) o.y >> 16, p3, s);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  return s << subShift;
}

struct TzCtx
{
  const int16_t* pat;       // pattern: shared memory (per-call jobs) or the current picture (frame search)
  int            patStride;
  const int16_t* refAtPU;   // global
  int            refStride, w, h, subShift;
  int            predQx, predQy, imvShift;
  double         lambda;
  TzSmem*        sm;
  int            staged;    // cStruct.subShiftMode == 1: xTZSearchHelp's staged SAD (tz_eval_staged)
};

template <class CTX>
__device__ __forceinline__ uint32_t tz_cost(const CTX& c, int x, int y, uint32_t sad)
{
  return sad + mv_cost(c.lambda, mv_bits_q(x * 4, y * 4, c.predQx, c.predQy, c.imvShift));
}

// Ordered argmin of a batch of n points (uniform across the cooperating warps).  Returns the winner's index in the
// batch and its cost, or -1 when no point beats `best` (strictly).
template <int WARPS, class PointFn>
__device__ __forceinline__ int tz_eval(const TzCtx& c, int n, uint32_t best, uint32_t& costOut, PointFn point)
{
  unsigned long long k = ~0ull;
  if (WARPS > 1)
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (threadIdx.x == 0) c.sm->key = ~0ull;
    __syncthreads();
    for (int i = warp; i < n; i += WARPS)
    {
      int x, y;
      point(i, x, y);
      const uint32_t sad = tz_warp_sad(c.pat, c.patStride, c.w, c.h, c.refAtPU + (ptrdiff_t) y * c.refStride + x, c.refStride, c.subShift);
      const unsigned long long ki = ((unsigned long long) tz_cost(c, x, y, sad) << 32) | (uint32_t) i;
      k = ki < k ? ki : k;
    }
    if (lane == 0 && k != ~0ull) atomicMin(&c.sm->key, k);
    __syncthreads();
    k = c.sm->key;
  }
  else
  {
    for (int i = 0; i < n; i++)
    {
      int x, y;
      point(i, x, y);
      const uint32_t sad = tz_warp_sad(c.pat, c.patStride, c.w, c.h, c.refAtPU + (ptrdiff_t) y * c.refStride + x, c.refStride, c.subShift);
      const unsigned long long ki = ((unsigned long long) tz_cost(c, x, y, sad) << 32) | (uint32_t) i;
      k = ki < k ? ki : k;
    }
  }
  costOut = (uint32_t) (k >> 32);
  return (k != ~0ull && costOut < best) ? (int) (uint32_t) k : -1;
}

// Row sums of |pattern - block| binned by the stage of xTZSearchHelp's staged SAD (subShiftMode 1, :340-391) that reads the
// row: with S = subShift, stage 0 holds the rows 0 mod 2^S, stage k >= 1 the rows 2^(S-k) mod 2^(S-k+1).  One warp; the
// sums are returned in every lane.
__device__ __forceinline__ void tz_warp_sad_stages(const int16_t* pat, int patStride, int w, int h, const int16_t* cur,
                                                   int refStride, int S, uint32_t (&sum)[5])
{
  const int lane = threadIdx.x & 31;
  const int q = w >> 2, units = q * h;
#pragma unroll
  for (int k = 0; k < 5; k++) sum[k] = 0;
  for (int u = lane; u < units; u += 32)
  {
    const int      r = u / q, c = (u % q) << 2;
    const int      low   = r & ((1 << S) - 1);
    const int      stage = low == 0 ? 0 : S - (__ffs(low) - 1);
    const uint2    o = *reinterpret_cast<const uint2*>(pat + r * patStride + c);
    const int16_t* p = cur + (ptrdiff_t) r * refStride + c;
    uint32_t       v = 0;
    v = __sad((int) (short) (o.x & 0xffffu), (int) p[0], v);
    v = __sad((int) o.x >> 16, (int) p[1], v);
    v = __sad((int) (short) (o.y & 0xffffu), (int) p[2], v);
    v = __sad((int) o.y >> 16, (int) p[3], v);
#pragma unroll
    for (int k = 0; k < 5; k++) sum[k] += stage == k ? v : 0u;
  }
#pragma unroll
  for (int k = 0; k < 5; k++)
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], m);
}

// A batch of probes under the staged SAD of subShiftMode 1.  The reference accepts a probe only if the scaled-up partial
// sums of every stage pass against the best cost of that moment: (R0 << S) + bits < best, then ((R0 + .. + Rk) << (S-k))
// + bits <= best for the stages in between, and the exact cost F < best at the end.  These are estimates, so a probe is
// not characterised by its cost alone but by the pair (gate, F): it is accepted iff best > gate, gate = max(first
// estimate, the later ones - 1, F), and then best becomes F.  Gates and costs of the batch are computed in parallel —
// a probe whose first estimate already fails against the best cost at the start of the batch (the best only falls) stops
// there, which is the selective search's own saving — and one thread replays the sequential acceptance over them.
// gate and exact cost of one staged probe whose first estimate e0 = (R0 << S) + bits passed (one warp)
__device__ __forceinline__ void tz_staged_gate(const TzCtx& c, const int16_t* cur, uint32_t bits, uint32_t e0, uint32_t& gate,
                                               uint32_t& full)
{
  const int S = c.subShift;
  uint32_t  r[5];
  tz_warp_sad_stages(c.pat, c.patStride, c.w, c.h, cur, c.refStride, S, r);
  uint32_t part = r[0];
  gate          = e0;
  full          = e0;   // S == 0: the first sum is the exact one
  for (int k = 1; k <= S; k++)
  {
    part += r[k];
    const uint32_t e = (part << (S - k)) + bits;
    if (k == S)
      full = e;
    else if (e > 0 && e - 1 > gate)
      gate = e - 1;
  }
  if (full > gate) gate = full;
}

template <int WARPS, class PointFn>
__device__ __forceinline__ int tz_eval_staged(const TzCtx& c, int n, uint32_t best, uint32_t& costOut, PointFn point)
{
  const int S = c.subShift;
  if (WARPS == 1)
  {
    // one warp per search (batched frame search): the warp walks the probes in order against the running best cost,
    // exactly the reference's flow
    uint32_t b   = best;
    int      win = -1;
    for (int i = 0; i < n; i++)
    {
      int x, y;
      point(i, x, y);
      const int16_t* cur  = c.refAtPU + (ptrdiff_t) y * c.refStride + x;
      const uint32_t bits = mv_cost(c.lambda, mv_bits_q(x * 4, y * 4, c.predQx, c.predQy, c.imvShift));
      const uint32_t e0   = tz_warp_sad(c.pat, c.patStride, c.w, c.h, cur, c.refStride, S) + bits;
      if (e0 < b)
      {
        uint32_t gate, full;
        tz_staged_gate(c, cur, bits, e0, gate, full);
        if (b > gate)
        {
          b   = full;
          win = i;
        }
      }
    }
    costOut = b;
    return win;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();   // the previous batch's key has been read by everybody
  for (int i = warp; i < n; i += WARPS)
  {
    int x, y;
    point(i, x, y);
    const int16_t* cur  = c.refAtPU + (ptrdiff_t) y * c.refStride + x;
    const uint32_t bits = mv_cost(c.lambda, mv_bits_q(x * 4, y * 4, c.predQx, c.predQy, c.imvShift));
    uint32_t       gate = 0xffffffffu, full = 0xffffffffu;
    const uint32_t e0   = tz_warp_sad(c.pat, c.patStride, c.w, c.h, cur, c.refStride, S) + bits;   // (R0 << S) + bits
    if (e0 < best) tz_staged_gate(c, cur, bits, e0, gate, full);
    if (lane == 0)
    {
      c.sm->gate[i] = gate;
      c.sm->full[i] = full;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    uint32_t b = best, win = 0xffffffffu;
    for (int i = 0; i < n; i++)
    {
      const uint32_t g = c.sm->gate[i];
      if (b > g)
      {
        b   = c.sm->full[i];
        win = (uint32_t) i;
      }
    }
    c.sm->key = ((unsigned long long) b << 32) | win;
  }
  __syncthreads();
  const unsigned long long k = c.sm->key;
  costOut = (uint32_t) (k >> 32);
  return (int) (uint32_t) k;   // 0xffffffff -> -1
}

// evaluator of the generic path: WARPS warps per search, any block size, pattern and reference read from memory
template <int WARPS>
struct TzEvalWarps
{
  using Ctx = TzCtx;
  static constexpr int kBatch = kTzChunk;   // largest batch one eval call may take
  template <class PointFn>
  static __device__ __forceinline__ int eval(const Ctx& c, int n, uint32_t best, uint32_t& costOut, PointFn point)
  {
    if (c.staged) return tz_eval_staged<WARPS>(c, n, best, costOut, point);
    return tz_eval<WARPS>(c, n, best, costOut, point);
  }
  // distFunc as it stands (sub-sampled rows scaled up), first strict minimum: the history seeds
  template <class PointFn>
  static __device__ __forceinline__ int eval_plain(const Ctx& c, int n, uint32_t best, uint32_t& costOut, PointFn point)
  {
    return tz_eval<WARPS>(c, n, best, costOut, point);
  }
};

// Evaluator of the batched frame search: one warp per search, square CU of SIZE samples with row sub-sampling SS known
// at compile time.  The 32 lanes lie ALONG the rows of a probe — two samples (one 32-bit word) per lane, LX lanes per
// row, RPI = 32 / LX rows per load instruction — so that a load touches as few 128-byte lines as the block has rows
// (the probes come straight from L1/L2, whose cost is one wavefront per distinct line and instruction).  A probe at
// an odd x takes the second half of its word pair from the neighbouring lane.  K probes are in flight at a time; their
// per-lane partial sums are reduced with a butterfly that halves the number of live values at every stage (about one
// shuffle per probe) and leaves probe (i0 + id(lane)) in its lane, so the K rate terms are computed in parallel too.
template <int SIZE, int SS, int COOP = 1>   // COOP warps share the rows of every probe (large CUs: one CTA per search)
struct TzEvalTile
{
  static constexpr int  WPR    = SIZE / 2;                  // 32-bit words per row
  static constexpr int  LX     = WPR < 32 ? WPR : 32;       // lanes along x
  static constexpr int  CB     = WPR / LX;                  // column blocks per row (128 wide: 2)
  static constexpr int  RPI    = 32 / LX;                   // rows per load instruction
  static constexpr int  ROWS   = SIZE >> SS;                // sampled rows
  static constexpr int  RITERS = ROWS / RPI;                // row steps per probe
  static constexpr int  ITERS  = RITERS * CB;               // load instructions per probe
  static constexpr bool PATREG = ITERS <= 16;               // the lane's pattern samples stay in registers
  static constexpr int  K      = SIZE == 8 ? 16 : (SIZE == 16 ? 8 : 1);
  struct Ctx
  {
    int            pat[PATREG ? ITERS * 2 : 2];
    const int16_t* patLane;     // pattern (current picture, rows 16-byte aligned) at this lane's first sample
    const int16_t* refLane;     // reference plane at the PU position + this lane's first sample
    int            patStep, refStep;   // element offset between two row steps of a lane: (RPI << SS) * stride
    int            refStride;
    int            predQx, predQy, imvShift;
    double         lambda;
    uint32_t*      partial;     // COOP > 1: shared memory, [2][COOP] per-warp partial sums (double-buffered)
    mutable int    parity;
  };

  static __device__ __forceinline__ void init(Ctx& c, const int16_t* patAtPU, int patStride, const int16_t* refAtPU, int refStride)
  {
    const int lane = threadIdx.x & 31, ri = lane / LX, lx = lane % LX;
    c.patLane   = patAtPU + (ri << SS) * patStride + lx * 2;
    c.refLane   = refAtPU + (ri << SS) * refStride + lx * 2;
    c.patStep   = (RPI << SS) * patStride;
    c.refStep   = (RPI << SS) * refStride;
    c.refStride = refStride;
    c.parity    = 0;
    if (PATREG)
    {
#pragma unroll
      for (int it = 0; it < ITERS; it++)
      {
        const uint32_t w  = *reinterpret_cast<const uint32_t*>(c.patLane + (it / CB) * c.patStep + (it % CB) * 64);
        c.pat[2 * it]     = (int) (w & 0xffffu);
        c.pat[2 * it + 1] = (int) (w >> 16);
      }
    }
  }

  // this lane's share of SAD(pattern, block at element offset `off` from the PU position)
  static __device__ __forceinline__ uint32_t lane_sad(const Ctx& c, int off)
  {
    const int16_t* q    = c.refLane + off;
    const unsigned sh   = (reinterpret_cast<uintptr_t>(q) & 2) ? 16u : 0u;   // odd x (strides are even: same for all rows)
    const bool     last = (threadIdx.x & 31) % LX == LX - 1;
    uint32_t       acc  = 0;
    if (PATREG)
    {
#pragma unroll
      for (int it = 0; it < ITERS; it++)
      {
        const uint32_t* wa = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(q + (it / CB) * c.refStep + (it % CB) * 64) & ~(uintptr_t) 3);
        const uint32_t  w0 = __ldg(wa);
        uint32_t        w1 = __shfl_down_sync(0xffffffffu, w0, 1);
        if (last && sh) w1 = __ldg(wa + 1);
        const uint32_t a = __funnelshift_r(w0, w1, sh);
        acc = __sad(c.pat[2 * it], (int) (a & 0xffffu), acc);
        acc = __sad(c.pat[2 * it + 1], (int) (a >> 16), acc);
      }
    }
    else
    {
      const int       w0i = COOP > 1 ? (int) (threadIdx.x >> 5) : 0;
      const int16_t*  pr  = q + w0i * c.refStep;
      const int16_t*  pp  = c.patLane + w0i * c.patStep;
#pragma unroll 4
      for (int rit = w0i; rit < RITERS; rit += COOP, pr += COOP * c.refStep, pp += COOP * c.patStep)
#pragma unroll
        for (int cb = 0; cb < CB; cb++)
        {
          const uint32_t* wa = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(pr + cb * 64) & ~(uintptr_t) 3);
          const uint32_t  w0 = __ldg(wa);
          uint32_t        w1 = __shfl_down_sync(0xffffffffu, w0, 1);
          if (last && sh) w1 = __ldg(wa + 1);
          const uint32_t a = __funnelshift_r(w0, w1, sh);
          const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(pp + cb * 64));
          acc = __sad((int) (w & 0xffffu), (int) (a & 0xffffu), acc);
          acc = __sad((int) (w >> 16), (int) (a >> 16), acc);
        }
    }
    return acc;
  }

  // Sums s[q] over the 32 lanes for all q < KK at once; afterwards the lane holds the total of probe id(lane):
  // stage with lane-mask m and `half` live values: a lane whose bit m is set keeps the upper half, sends the lower.
  template <int KK>
  static __device__ __forceinline__ uint32_t butterfly(uint32_t (&s)[KK], int lane, int& id)
  {
    int m = 16;
    id    = 0;
#pragma unroll
    for (int half = KK / 2; half >= 1; half >>= 1, m >>= 1)
    {
      const bool up = (lane & m) != 0;
      if (up) id += half;
#pragma unroll
      for (int j = 0; j < half; j++)
      {
        const uint32_t send = up ? s[j] : s[j + half];
        const uint32_t keep = up ? s[j + half] : s[j];
        s[j]                = keep + __shfl_xor_sync(0xffffffffu, send, m);
      }
    }
#pragma unroll
    for (; m >= 1; m >>= 1) s[0] += __shfl_xor_sync(0xffffffffu, s[0], m);
    return s[0];
  }

  // KK probes i0 .. i0+KK-1 (all valid) in flight; updates the lane's running (cost, index) minimum
  template <int KK, class PointFn>
  static __device__ __forceinline__ void phase(const Ctx& c, int i0, uint32_t& bestCost, uint32_t& bestIdx, PointFn point)
  {
    const int lane = threadIdx.x & 31;
    uint32_t  s[KK];
#pragma unroll
    for (int q = 0; q < KK; q++)
    {
      int x, y;
      point(i0 + q, x, y);
      s[q] = lane_sad(c, y * c.refStride + x);
    }
    int      id;
    uint32_t sad = butterfly<KK>(s, lane, id);
    if (COOP > 1)
    {
      // every lane holds its warp's share of the probe; one barrier per probe (the buffer written now is read before
      // the barrier of the next probe, which every warp passes before writing it again)
      uint32_t* buf = c.partial + c.parity * COOP;
      c.parity ^= 1;
      if (lane == 0) buf[threadIdx.x >> 5] = sad;
      __syncthreads();
      sad = 0;
#pragma unroll
      for (int w = 0; w < COOP; w++) sad += buf[w];
    }
    int x, y;
    point(i0 + id, x, y);
    const uint32_t cost = tz_cost(c, x, y, sad << SS);
    if (cost < bestCost)   // probes of later phases have larger indices: strict
    {
      bestCost = cost;
      bestIdx  = (uint32_t) (i0 + id);
    }
  }

  template <class PointFn>
  static __device__ __forceinline__ int eval_plain(const Ctx& c, int n, uint32_t best, uint32_t& costOut, PointFn point)
  {
    return eval(c, n, best, costOut, point);
  }
  template <class PointFn>
  static __device__ __forceinline__ int eval(const Ctx& c, int n, uint32_t best, uint32_t& costOut, PointFn point)
  {
    uint32_t bc = 0xffffffffu, bi = 0xffffffffu;
    int      i0 = 0;
    for (; i0 + K <= n; i0 += K) phase<K>(c, i0, bc, bi, point);
    if (K >= 16 && n - i0 >= 8) { phase<8>(c, i0, bc, bi, point); i0 += 8; }
    if (K >= 8 && n - i0 >= 4) { phase<4>(c, i0, bc, bi, point); i0 += 4; }
    if (K >= 4 && n - i0 >= 2) { phase<2>(c, i0, bc, bi, point); i0 += 2; }
    if (K >= 2 && n - i0 >= 1) { phase<1>(c, i0, bc, bi, point); i0 += 1; }
    // first strict minimum in probe order over the lanes: smallest cost, then smallest index among its holders
    const uint32_t mc = __reduce_min_sync(0xffffffffu, bc);
    const uint32_t mi = __reduce_min_sync(0xffffffffu, bc == mc ? bi : 0xffffffffu);
    costOut           = mc;
    return (mi != 0xffffffffu && mc < best) ? (int) mi : -1;
  }
};

struct TzPoints
{
  short         x[kTzMaxPoints], y[kTzMaxPoints];
  unsigned char  pnr[kTzMaxPoints];
  unsigned short dist[kTzMaxPoints];
  int           n;
};

// a coordinate that moved away from the start is checked against the window on its side, one equal to the start's is
// not: the rule every branch of xTZ8PointDiamondSearch follows
__device__ __forceinline__ void tz_add(TzPoints& p, const TzState& s, int sx, int sy, int dx, int dy, int pnr, int dist)
{
  const int x = sx + dx, y = sy + dy;
  if ((dx < 0 && x < s.l) || (dx > 0 && x > s.r) || (dy < 0 && y < s.t) || (dy > 0 && y > s.b)) return;
  p.x[p.n]    = (short) x;
  p.y[p.n]    = (short) y;
  p.pnr[p.n]  = (unsigned char) pnr;
  p.dist[p.n] = (unsigned short) dist;
  p.n++;
}

// xTZSearchHelp's state update for the winner of a batch
__device__ __forceinline__ void tz_update(TzState& s, uint32_t cost, int x, int y, int pnr, uint32_t dist)
{
  s.best  = cost;
  s.bx    = x;
  s.by    = y;
  s.dist  = dist;
  s.round = 0;
  s.pnr   = pnr;
}

template <class EV>
__device__ __forceinline__ void tz_run_points(const typename EV::Ctx& c, TzState& s, const TzPoints& p)
{
  uint32_t  cost;
  const int win = EV::eval(c, p.n, s.best, cost, [&](int i, int& x, int& y) { x = p.x[i]; y = p.y[i]; });
  if (win >= 0) tz_update(s, cost, p.x[win], p.y[win], p.pnr[win], p.dist[win]);
}

// xTZ8PointDiamondSearch (:503-705)
template <class EV>
__device__ inline void tz_diamond(const typename EV::Ctx& c, TzState& s, int sx, int sy, int d, bool corners)
{
  TzPoints p;
  p.n = 0;
  s.round += 1;
  if (d == 1)
  {
    if (corners)
    {
      if (sy - 1 >= s.t)
      {
        tz_add(p, s, sx, sy, -1, -1, 1, 1);
        tz_add(p, s, sx, sy, 0, -1, 2, 1);
        tz_add(p, s, sx, sy, 1, -1, 3, 1);
      }
    }
    else
      tz_add(p, s, sx, sy, 0, -1, 2, 1);
    tz_add(p, s, sx, sy, -1, 0, 4, 1);
    tz_add(p, s, sx, sy, 1, 0, 5, 1);
    if (corners)
    {
      if (sy + 1 <= s.b)
      {
        tz_add(p, s, sx, sy, -1, 1, 6, 1);
        tz_add(p, s, sx, sy, 0, 1, 7, 1);
        tz_add(p, s, sx, sy, 1, 1, 8, 1);
      }
    }
    else
      tz_add(p, s, sx, sy, 0, 1, 7, 1);
  }
  else if (d <= 8)
  {
    const int h = d >> 1;
    tz_add(p, s, sx, sy, 0, -d, 2, d);
    tz_add(p, s, sx, sy, -h, -h, 1, h);
    tz_add(p, s, sx, sy, h, -h, 3, h);
    tz_add(p, s, sx, sy, -d, 0, 4, d);
    tz_add(p, s, sx, sy, d, 0, 5, d);
    tz_add(p, s, sx, sy, -h, h, 6, h);
    tz_add(p, s, sx, sy, h, h, 8, h);
    tz_add(p, s, sx, sy, 0, d, 7, d);
  }
  else
  {
    tz_add(p, s, sx, sy, 0, -d, 0, d);
    tz_add(p, s, sx, sy, -d, 0, 0, d);
    tz_add(p, s, sx, sy, d, 0, 0, d);
    tz_add(p, s, sx, sy, 0, d, 0, d);
    for (int i = 1; i < 4; i++)
    {
      const int q = (d >> 2) * i;
      tz_add(p, s, sx, sy, -q, -d + q, 0, d);
      tz_add(p, s, sx, sy, q, -d + q, 0, d);
      tz_add(p, s, sx, sy, -q, d - q, 0, d);
      tz_add(p, s, sx, sy, q, d - q, 0, d);
    }
  }
  tz_run_points<EV>(c, s, p);
}

// xTZ2PointSearch (:420-446)
template <class EV>
__device__ inline void tz_two_points(const typename EV::Ctx& c, TzState& s)
{
  const int xo[2][9] = { { 0, -1, -1, 0, -1, +1, -1, -1, +1 }, { 0, 0, +1, +1, -1, +1, 0, +1, 0 } };
  const int yo[2][9] = { { 0, 0, -1, -1, +1, -1, 0, +1, 0 }, { 0, -1, -1, 0, -1, +1, +1, +1, +1 } };
  TzPoints  p;
  p.n = 0;
  for (int k = 0; k < 2; k++)
  {
    const int x = s.bx + xo[k][s.pnr], y = s.by + yo[k][s.pnr];
    if (x >= s.l && x <= s.r && y >= s.t && y <= s.b)
    {
      p.x[p.n]    = (short) x;
      p.y[p.n]    = (short) y;
      p.pnr[p.n]  = 0;
      p.dist[p.n] = 2;
      p.n++;
    }
  }
  tz_run_points<EV>(c, s, p);
}

// raster scan over [l,r] x [t,b] with step `win` (:3876-3898), in chunks of kTzChunk points
template <class EV>
__device__ inline void tz_raster(const typename EV::Ctx& c, TzState& s, int l, int r, int t, int b, int win)
{
  if (r < l || b < t) return;
  const int nx = (r - l) / win + 1, ny = (b - t) / win + 1, n = nx * ny;
  // g / nx as a multiply-high: exact for g < 2^17 and nx <= 1099 (checked exhaustively), far above any window here
  const uint32_t inv = 0xffffffffu / (uint32_t) nx + 1u;
  for (int base = 0; base < n; base += kTzChunk)
  {
    uint32_t  cost;
    const int m   = min(kTzChunk, n - base);
    const int idx = EV::eval(c, m, s.best, cost, [&](int i, int& x, int& y) {
      const int g = base + i, gy = (int) __umulhi((uint32_t) g, inv);
      x           = l + (g - gy * nx) * win;
      y           = t + gy * win;
    });
    if (idx >= 0)
    {
      const int g = base + idx, gy = g / nx;
      tz_update(s, cost, l + (g - gy * nx) * win, t + gy * win, 0, (uint32_t) win);
    }
  }
}

// clipMv + changePrecision(INTERNAL -> QUARTER) + divideByPowerOf2(2)  (:3682-3683)
__device__ __forceinline__ void tz_to_int(const DevTz& t, int& x, int& y)
{
  const int horMax = (t.picW + 8 - t.posX - 1) * 16, horMin = (-t.maxCuW - 8 - t.posX + 1) * 16;
  const int verMax = (t.picH + 8 - t.posY - 1) * 16, verMin = (-t.maxCuH - 8 - t.posY + 1) * 16;
  x = clampi(x, horMin, horMax);
  y = clampi(y, verMin, verMax);
  x = x >= 0 ? (x + 1) >> 2 : (x + 2) >> 2;
  y = y >= 0 ? (y + 1) >> 2 : (y + 2) >> 2;
  x = div_pow2_round(x, 2);
  y = div_pow2_round(y, 2);
}

// history MVs (xTZSearch :3734-3765, xTZSearchSelective :4040-4074): duplicates of an earlier entry are skipped; the
// distortion is distFunc's own, also under subShiftMode 1; only position and cost are updated
template <class EV>
__device__ inline void tz_seeds(const typename EV::Ctx& c, const DevTz& t, TzState& s)
{
  if (t.nSeeds <= 0) return;
  TzPoints  p;
  const int horMax = (t.picW + 8 - t.posX - 1) * 16, horMin = (-t.maxCuW - 8 - t.posX + 1) * 16;
  const int verMax = (t.picH + 8 - t.posY - 1) * 16, verMin = (-t.maxCuH - 8 - t.posY + 1) * 16;
  p.n = 0;
  for (int i = 0; i < t.nSeeds; i++)
  {
    int k = 0;
    for (; k < i; k++)
      if (t.seedX[k] == t.seedX[i] && t.seedY[k] == t.seedY[i]) break;
    if (k < i) continue;
    const int x = clampi(t.seedX[i], horMin, horMax), y = clampi(t.seedY[i], verMin, verMax);
    p.x[p.n] = (short) (x >= 0 ? (x + 7) >> 4 : (x + 8) >> 4);   // changePrecision(INTERNAL -> INT)
    p.y[p.n] = (short) (y >= 0 ? (y + 7) >> 4 : (y + 8) >> 4);
    p.n++;
  }
  uint32_t  cost;
  const int win = EV::eval_plain(c, p.n, s.best, cost, [&](int i, int& x, int& y) { x = p.x[i]; y = p.y[i]; });
  if (win >= 0)
  {
    s.best = cost;
    s.bx   = p.x[win];
    s.by   = p.y[win];
  }
}

// The whole xTZSearch.  Returns the best key (cost, position) to every thread of the cooperating warps.
template <class EV>
__device__ inline unsigned long long tz_search(const typename EV::Ctx& c, const DevTz& t)
{
  const int raster = t.fast ? 8 : 5;
  const int range  = t.searchRange;
  TzState   s;
  s.best = 0xffffffffu;
  s.bx = s.by = 0;
  s.dist = s.round = 0;
  s.pnr = 0;
  s.l = s.r = s.t = s.b = 0;

  int sx = t.startX, sy = t.startY;
  tz_to_int(t, sx, sy);
  {
    // start point, zero vector, 2Nx2N integer MV: each decision depends on the previous probe (:3695-3732)
    TzPoints p;
    p.n       = 1;
    p.x[0]    = (short) sx;
    p.y[0]    = (short) sy;
    p.pnr[0]  = 0;
    p.dist[0] = 0;
    tz_run_points<EV>(c, s, p);
    if (!t.fast && (sx != 0 || sy != 0) && (s.bx != 0 || s.by != 0))
    {
      p.x[0] = p.y[0] = 0;
      tz_run_points<EV>(c, s, p);
    }
    if (t.hasInt2Nx2N)
    {
      int ix = t.int2Nx2NX * 16, iy = t.int2Nx2NY * 16;
      tz_to_int(t, ix, iy);
      if ((sx != ix || sy != iy) && (ix != s.bx || iy != s.by))
      {
        p.x[0] = (short) ix;
        p.y[0] = (short) iy;
        tz_run_points<EV>(c, s, p);
      }
    }
  }
  tz_seeds<EV>(c, t, s);
  {
    // xSetSearchRange around the best start point (:3767-3772)
    const Window w = search_window(s.bx * 4, s.by * 4, t.posX, t.posY, t.picW, t.picH, t.maxCuW, range >> (t.fast ? 1 : 0));
    s.l = w.l;
    s.r = w.r;
    s.t = w.t;
    s.b = w.b;
  }
  sx = s.bx;
  sy = s.by;
  const bool bestIsZero = s.bx == 0 && s.by == 0;
  for (int d = 1; d <= range; d *= 2)   // first search (:3803-3818)
  {
    tz_diamond<EV>(c, s, sx, sy, d, t.extended != 0);
    if (t.firstSearchStop && s.round >= 3u) break;
  }
  if (t.extended && !bestIsZero)   // zero neighbourhood with half the range (:3841-3855)
    for (int d = 1; d <= (range >> 1); d *= 2) tz_diamond<EV>(c, s, 0, 0, d, false);
  if (s.dist == 1)   // :3858-3863
  {
    s.dist = 0;
    tz_two_points<EV>(c, s);
  }
  if (t.extended)   // adaptive raster (:3865-3885)
  {
    int win = raster, l = s.l, r = s.r, tt = s.t, b = s.b;
    if (!((int) s.dist >= raster))
    {
      win++;
      l /= 2;
      r /= 2;
      tt /= 2;
      b /= 2;
    }
    s.dist = (uint32_t) win;
    tz_raster<EV>(c, s, l, r, tt, b, win);
  }
  else if ((int) s.dist >= raster)   // :3886-3899
  {
    s.dist = (uint32_t) raster;
    tz_raster<EV>(c, s, s.l, s.r, s.t, s.b, raster);
  }
  while (s.dist > 0)   // star refinement (:3932-3967)
  {
    sx     = s.bx;
    sy     = s.by;
    s.dist = 0;
    s.pnr  = 0;
    for (int d = 1; d < range + 1; d *= 2)
    {
      tz_diamond<EV>(c, s, sx, sy, d, t.extended != 0);
      if (t.fast && s.round >= 2u) break;
    }
    if (s.dist == 1)
    {
      s.dist = 0;
      if (s.pnr != 0) tz_two_points<EV>(c, s);
    }
  }
  return make_key(s.best, s.bx, s.by);
}

// The whole xTZSearchSelective (:3979-4170, FastSearch=2 / MESEARCH_SELECTIVE; no hash ME).  Start points, then a grid of
// step 4 within +-range/4 of the best start point, every grid point with its diamonds of distance 1 and 2 — 13 fixed
// probes per grid point, one batch — then either the exhaustive scan of the window (the best moved more than 8 samples
// away) or the star refinement without a stop criterion.
template <class EV>
__device__ inline unsigned long long tz_search_selective(const typename EV::Ctx& c, const DevTz& t)
{
  const int range = t.searchRange, rangeInitial = t.searchRange >> 2;
  TzState   s;
  s.best = 0xffffffffu;
  s.bx = s.by = 0;
  s.dist = s.round = 0;
  s.pnr = 0;
  s.l = s.r = s.t = s.b = 0;
  {
    // median predictor, zero vector, 2Nx2N integer MV: unconditional here (:4017-4038); a repeated position costs the
    // same and cannot win again
    TzPoints p;
    int      sx = t.startX, sy = t.startY;
    tz_to_int(t, sx, sy);
    p.n = 2;
    p.x[0] = (short) sx;
    p.y[0] = (short) sy;
    p.x[1] = p.y[1] = 0;
    if (t.hasInt2Nx2N)
    {
      int ix = t.int2Nx2NX * 16, iy = t.int2Nx2NY * 16;
      tz_to_int(t, ix, iy);
      p.x[2] = (short) ix;
      p.y[2] = (short) iy;
      p.n    = 3;
    }
    for (int i = 0; i < p.n; i++)
    {
      p.pnr[i]  = 0;
      p.dist[i] = 0;
    }
    tz_run_points<EV>(c, s, p);
  }
  tz_seeds<EV>(c, t, s);
  {
    // :4076-4081 — the reference shifts the integer position by 2 (not by MV_FRACTIONAL_BITS_INTERNAL) before
    // xSetSearchRange reads it as a 1/16-sample vector: the window is centred on a quarter of the best start point
    const Window w = search_window(s.bx, s.by, t.posX, t.posY, t.picW, t.picH, t.maxCuW, range);
    s.l = w.l;
    s.r = w.r;
    s.t = w.t;
    s.b = w.b;
  }
  const int bx0 = s.bx, by0 = s.by;
  {
    // initial search (:4104-4120)
    const int l = max(bx0 - rangeInitial, s.l), r = min(bx0 + rangeInitial, s.r);
    const int tt = max(by0 - rangeInitial, s.t), b = min(by0 + rangeInitial, s.b);
    if (l <= r && tt <= b)
    {
      // probe order of one grid point: itself, xTZ8PointDiamondSearch(1) = up, left, right, down, then (2) = the eight
      // points of tz_diamond's d <= 8 branch; packed as (dx + 2) | (dy + 2) << 3 | dist << 6
      constexpr int kPer = 13;
      constexpr int kGridPerBatch = EV::kBatch / kPer;
      const int nx = (r - l) / 4 + 1, ny = (b - tt) / 4 + 1, nGrid = nx * ny;
      auto      offset = [](int k, int& dx, int& dy, int& dist) {
        const int dxs[kPer] = { 0, 0, -1, 1, 0, 0, -1, 1, -2, 2, -1, 1, 0 };
        const int dys[kPer] = { 0, -1, 0, 0, 1, -2, -1, -1, 0, 0, 1, 1, 2 };
        const int dst[kPer] = { 0, 1, 1, 1, 1, 2, 1, 1, 2, 2, 1, 1, 2 };
        dx   = dxs[k];
        dy   = dys[k];
        dist = dst[k];
      };
      for (int g0 = 0; g0 < nGrid; g0 += kGridPerBatch)
      {
        const int m     = min(kGridPerBatch, nGrid - g0) * kPer;
        auto      point = [&](int i, int& x, int& y) {
          const int g = g0 + i / kPer, k = i % kPer, gy = g / nx, gx = g - gy * nx;
          int       dx, dy, dist;
          offset(k, dx, dy, dist);
          const int cx = l + 4 * gx, cy = tt + 4 * gy;
          x = cx + dx;
          y = cy + dy;
          // a point outside the window is not probed (tz_add's rule); here it repeats the grid point, which cannot win
          if ((dx < 0 && x < s.l) || (dx > 0 && x > s.r) || (dy < 0 && y < s.t) || (dy > 0 && y > s.b))
          {
            x = cx;
            y = cy;
          }
        };
        uint32_t  cost;
        const int win = EV::eval(c, m, s.best, cost, point);
        if (win >= 0)
        {
          int x, y, dx, dy, dist;
          point(win, x, y);
          offset(win % kPer, dx, dy, dist);
          tz_update(s, cost, x, y, 0, (uint32_t) dist);
        }
      }
    }
  }
  if (abs(s.bx - bx0) > 8 || abs(s.by - by0) > 8)   // :4122-4134
    tz_raster<EV>(c, s, s.l, s.r, s.t, s.b, 1);
  else
    while (s.dist > 0)   // :4136-4165
    {
      const int sx = s.bx, sy = s.by;
      s.dist = 0;
      s.pnr  = 0;
      for (int d = 1; d < range + 1; d *= 2) tz_diamond<EV>(c, s, sx, sy, d, false);
      if (s.dist == 1)
      {
        s.dist = 0;
        if (s.pnr != 0) tz_two_points<EV>(c, s);
      }
    }
  return make_key(s.best, s.bx, s.by);
}

}   // namespace vtmme
