// Batched TZ search of the frame path (vtmme_search_frames with fastSearch 1 or 3; 2 = the selective search, see
// me_tz_frame_selective_kernel): every grid-aligned square CU of
// every picture pair runs InterSearch::xTZSearch (EncoderLib/InterSearch.cpp:3640-3974, me_tz.cuh) from its own
// predictor and leaves the (cost, position) key the fractional refinement (me_frac.cu) starts from — the same
// hand-over as the full search's tree kernels.
//
// One warp per CU, four CUs per CTA, nothing but warp-level synchronisation: the searches are chains of dependent
// probe batches, so throughput comes from the number of searches in flight.  A probe occupies as many lanes as the
// CU has sampled rows (TzEvalTile in me_tz.cuh: four 8x8 probes per warp at a time), the lane's pattern row stays in
// registers, the probes are read from the reference picture through L1/L2 — a search touches a few hundred
// scattered blocks of a window that neighbouring CUs share.  Largest CUs are scheduled first.
#include <cstdlib>

#include "me_tz.cuh"

namespace vtmme {

namespace {

__host__ __device__ constexpr int coop_warps(int size) { return size >= 64 ? 4 : 1; }   // warps per search

template <int SIZE, int SS>
__device__ __forceinline__ unsigned long long tz_frame_cu(const TzFrameParams& p, const DevPic& cur, const DevPic& ref, int x, int y,
                                                          short2 pr, const DevTz& t, uint32_t* partial)
{
  using EV = TzEvalTile<SIZE, SS, coop_warps(SIZE)>;
  typename EV::Ctx c;
  c.partial  = partial;
  c.predQx   = pr.x;
  c.predQy   = pr.y;
  c.imvShift = p.imvShift;
  c.lambda   = p.lambda;
  EV::init(c, cur.origin + (ptrdiff_t) y * cur.stride + x, cur.stride, ref.origin + (ptrdiff_t) y * ref.stride + x, ref.stride);
  return tz_search<EV>(c, t);
}

// one launch per CU level (its own register budget and occupancy); a CTA carries four searches of CUs up to 32x32 or
// one search of a larger CU, whose rows its four warps share
template <int SIZE, int SS>
__global__ void __launch_bounds__(128) me_tz_frame_kernel(TzFrameParams p, int level)
{
  __shared__ uint32_t s_partial[8];
  constexpr int perCta = 4 / coop_warps(SIZE);
  const int warp = threadIdx.x >> 5;
  const int nCU  = p.g.off[5], nLevel = p.g.nx[level] * p.g.ny[level];
  const int li = blockIdx.x * perCta + (perCta > 1 ? warp : 0), pair = blockIdx.y;
  if (li >= nLevel) return;
  const int cu = p.g.off[level] + li;
  const int x = (li % p.g.nx[level]) * SIZE, y = (li / p.g.nx[level]) * SIZE;
  const DevPic cur = p.cur[pair], ref = p.ref[pair];
  short2       pr  = make_short2(0, 0);
  if (p.predQ) pr = p.predQ[(size_t) pair * nCU + cu];

  DevTz t;
  t.startX = pr.x * 4;   // rcMv = rcMvPred (InterSearch.cpp:3451), quarter-pel -> 1/16
  t.startY = pr.y * 4;
  t.hasInt2Nx2N = 0;
  t.int2Nx2NX = t.int2Nx2NY = 0;
  t.nSeeds          = 0;
  t.searchRange     = p.sr;
  t.extended        = p.extended;
  t.fast            = 0;
  t.firstSearchStop = p.firstSearchStop;
  t.posX            = x;
  t.posY            = y;
  t.picW            = p.g.picW;
  t.picH            = p.g.picH;
  t.maxCuW = t.maxCuH = p.ctu;
  const unsigned long long key = tz_frame_cu<SIZE, SS>(p, cur, ref, x, y, pr, t, s_partial);
  if (threadIdx.x == 0 || (perCta > 1 && (threadIdx.x & 31) == 0)) p.keys[(size_t) pair * nCU + cu] = key;
}

// FastSearch=2: xTZSearchSelective (InterSearch.cpp:3979-4170) per CU, one warp per search on the generic evaluator
// (pattern read from the current picture); subShiftMode 1 = the staged SAD of xTZSearchHelp (:340-391)
__global__ void __launch_bounds__(128) me_tz_frame_selective_kernel(TzFrameParams p, int level)
{
  const int size = 8 << level;
  const int nCU  = p.g.off[5], nLevel = p.g.nx[level] * p.g.ny[level];
  const int li = blockIdx.x * 4 + (threadIdx.x >> 5), pair = blockIdx.y;
  if (li >= nLevel) return;
  const int cu = p.g.off[level] + li;
  const int x = (li % p.g.nx[level]) * size, y = (li / p.g.nx[level]) * size;
  const DevPic cur = p.cur[pair], ref = p.ref[pair];
  short2       pr  = make_short2(0, 0);
  if (p.predQ) pr = p.predQ[(size_t) pair * nCU + cu];

  DevTz t;
  t.startX = pr.x * 4;
  t.startY = pr.y * 4;
  t.hasInt2Nx2N = 0;
  t.int2Nx2NX = t.int2Nx2NY = 0;
  t.nSeeds          = 0;
  t.searchRange     = p.sr;
  t.extended        = 0;
  t.fast            = 0;
  t.firstSearchStop = 0;
  t.posX            = x;
  t.posY            = y;
  t.picW            = p.g.picW;
  t.picH            = p.g.picH;
  t.maxCuW = t.maxCuH = p.ctu;
  t.selective       = 1;
  t.staged          = p.subShiftMode == 1;

  TzCtx c;
  c.pat       = cur.origin + (ptrdiff_t) y * cur.stride + x;   // rows 16-byte aligned (x is a multiple of 8 samples)
  c.patStride = cur.stride;
  c.refAtPU   = ref.origin + (ptrdiff_t) y * ref.stride + x;
  c.refStride = ref.stride;
  c.w = c.h   = size;
  // DistParam::subShift (RdCost.cpp:289-323): mode 1 by height, mode 2 for 16..64 rows
  c.subShift  = p.subShiftMode == 1 ? (size > 32 ? 4 : (size > 16 ? 3 : (size > 8 ? 2 : 1)))
                                    : (p.subShiftMode == 2 && size > 8 && size <= 64 ? 1 : 0);
  c.predQx    = pr.x;
  c.predQy    = pr.y;
  c.imvShift  = p.imvShift;
  c.lambda    = p.lambda;
  c.sm        = nullptr;
  c.staged    = t.staged;
  const unsigned long long key = tz_search_selective<TzEvalWarps<1>>(c, t);
  if ((threadIdx.x & 31) == 0) p.keys[(size_t) pair * nCU + cu] = key;
}

template <int SIZE, int SS>
cudaError_t launch_level(const TzFrameParams& p, int level, int nPairs, cudaStream_t st)
{
  const int n = p.g.nx[level] * p.g.ny[level];
  if (n == 0) return cudaSuccess;
  constexpr int perCta = 4 / coop_warps(SIZE);
  dim3 grid((n + perCta - 1) / perCta, nPairs, 1);
  me_tz_frame_kernel<SIZE, SS><<<grid, 128, 0, st>>>(p, level);
  return cudaGetLastError();
}

}   // namespace

// DistParam::subShift of the integer search is 1 for CUs of 16..64 rows when subShiftMode is 2 (RdCost.cpp:310-316).
// Largest CUs first.  Returns the number of kernels launched in *launches.
cudaError_t launch_tz_frame(const TzFrameParams& p, int nPairs, int predSpread, cudaStream_t st, int* launches)
{
  const bool  ss = p.subShiftMode == 2;
  cudaError_t e;
  // levels 8x8 .. 64x64: windows staged in shared memory (me_tz_smem.cu) when they fit; VTMME_TZ_VARIANT=global keeps the
  // kernels below (probes read from the picture) for all levels
  static const bool useSmem = [] { const char* v = getenv("VTMME_TZ_VARIANT"); return !(v && v[0] == 'g'); }();
  bool viaSmem[4] = { false, false, false, false };
  if (p.selective)
  {
    for (int l = 4; l >= 0; l--)
    {
      const int n = p.g.nx[l] * p.g.ny[l];
      if (n == 0) continue;
      me_tz_frame_selective_kernel<<<dim3((n + 3) / 4, nPairs, 1), 128, 0, st>>>(p, l);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      *launches += 1;
    }
    return cudaSuccess;
  }
  if ((e = launch_level<128, 0>(p, 4, nPairs, st)) != cudaSuccess) return e;
  // measured per level (32 pairs of 1080p, SR=64, profiles/r02k_tz_levels.md): staged windows win at 8x8, 16x16 and 64x64
  // (38.7 / 22.9 / 19.3 ms against 46.5 / 26.5 / 26.2), the picture-read kernel at 32x32 (18.0 against 21.7)
  if (useSmem)
    for (int l = 3; l >= 0; l--)
      if (l != 2 && (e = launch_tz_frame_smem(p, l, nPairs, predSpread, st, &viaSmem[l])) != cudaSuccess) return e;
  if (!viaSmem[3] && (e = ss ? launch_level<64, 1>(p, 3, nPairs, st) : launch_level<64, 0>(p, 3, nPairs, st)) != cudaSuccess) return e;
  if (!viaSmem[2] && (e = ss ? launch_level<32, 1>(p, 2, nPairs, st) : launch_level<32, 0>(p, 2, nPairs, st)) != cudaSuccess) return e;
  if (!viaSmem[1] && (e = ss ? launch_level<16, 1>(p, 1, nPairs, st) : launch_level<16, 0>(p, 1, nPairs, st)) != cudaSuccess) return e;
  if (!viaSmem[0] && (e = launch_level<8, 0>(p, 0, nPairs, st)) != cudaSuccess) return e;
  for (int l = 0; l < 5; l++) *launches += p.g.nx[l] * p.g.ny[l] > 0;
  return cudaSuccess;
}

}   // namespace vtmme
