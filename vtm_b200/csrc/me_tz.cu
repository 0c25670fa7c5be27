// Batched TZ search of the frame path (vtmme_search_frames with fastSearch 1 or 3): every grid-aligned square CU of
// every picture pair runs InterSearch::xTZSearch (EncoderLib/InterSearch.cpp:3640-3974, me_tz.cuh) from its own
// predictor and leaves the (cost, position) key the fractional refinement (me_frac.cu) starts from — the same
// hand-over as the full search's tree kernels.
//
// One warp per CU, four CUs per CTA, nothing but warp-level synchronisation: the searches are chains of dependent
// probe batches, so throughput comes from the number of searches in flight.  A probe occupies as many lanes as the
// CU has sampled rows (TzEvalTile in me_tz.cuh: four 8x8 probes per warp at a time), the lane's pattern row stays in
// registers, the probes are read from the reference picture through L1/L2 — a search touches a few hundred
// scattered blocks of a window that neighbouring CUs share.  Largest CUs are scheduled first.
#include "me_tz.cuh"

namespace vtmme {

namespace {

template <int SIZE, int SS>
__device__ __forceinline__ unsigned long long tz_frame_cu(const TzFrameParams& p, const DevPic& cur, const DevPic& ref, int x, int y,
                                                          short2 pr, const DevTz& t)
{
  using EV = TzEvalTile<SIZE, SS>;
  typename EV::Ctx c;
  c.patPtr    = cur.origin + (ptrdiff_t) y * cur.stride + x;
  c.patStride = cur.stride;
  c.refAtPU   = ref.origin + (ptrdiff_t) y * ref.stride + x;
  c.refStride = ref.stride;
  c.predQx    = pr.x;
  c.predQy    = pr.y;
  c.imvShift  = p.imvShift;
  c.lambda    = p.lambda;
  EV::load_pattern(c);
  return tz_search<EV>(c, t);
}

__global__ void __launch_bounds__(128) me_tz_frame_kernel(TzFrameParams p)
{
  const int warp = threadIdx.x >> 5;
  const int nCU  = p.g.off[5];
  const int item = blockIdx.x * 4 + warp, pair = blockIdx.y;
  if (item >= nCU) return;
  const int cu = nCU - 1 - item;   // level-major order, reversed: 128x128 CUs first
  int       level = 0;
#pragma unroll
  for (int l = 1; l < 5; l++)
    if (cu >= p.g.off[l]) level = l;
  const int size = 8 << level, li = cu - p.g.off[level];
  const int x = (li % p.g.nx[level]) * size, y = (li / p.g.nx[level]) * size;
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
  // DistParam::subShift of the integer search: 1 for CUs of 16..64 rows when subShiftMode is 2 (RdCost.cpp:310-316)
  const bool ss = p.subShiftMode == 2;
  unsigned long long key;
  switch (level)
  {
  case 0: key = tz_frame_cu<8, 0>(p, cur, ref, x, y, pr, t); break;
  case 1: key = ss ? tz_frame_cu<16, 1>(p, cur, ref, x, y, pr, t) : tz_frame_cu<16, 0>(p, cur, ref, x, y, pr, t); break;
  case 2: key = ss ? tz_frame_cu<32, 1>(p, cur, ref, x, y, pr, t) : tz_frame_cu<32, 0>(p, cur, ref, x, y, pr, t); break;
  case 3: key = ss ? tz_frame_cu<64, 1>(p, cur, ref, x, y, pr, t) : tz_frame_cu<64, 0>(p, cur, ref, x, y, pr, t); break;
  default: key = tz_frame_cu<128, 0>(p, cur, ref, x, y, pr, t); break;
  }
  if ((threadIdx.x & 31) == 0) p.keys[(size_t) pair * nCU + cu] = key;
}

}   // namespace

cudaError_t launch_tz_frame(const TzFrameParams& p, int nPairs, cudaStream_t st)
{
  dim3 grid((p.g.off[5] + 3) / 4, nPairs, 1);
  me_tz_frame_kernel<<<grid, 128, 0, st>>>(p);
  return cudaGetLastError();
}

}   // namespace vtmme
