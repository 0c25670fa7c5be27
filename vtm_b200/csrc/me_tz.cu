// Batched TZ search of the frame path (vtmme_search_frames with fastSearch 1 or 3): every grid-aligned square CU of
// every picture pair runs InterSearch::xTZSearch (EncoderLib/InterSearch.cpp:3640-3974, me_tz.cuh) from its own
// predictor and leaves the (cost, position) key the fractional refinement (me_frac.cu) starts from — the same
// hand-over as the full search's tree kernels.
//
// One warp per CU, four CUs per CTA, nothing but warp-level synchronisation: the searches are short chains of
// dependent probe batches, so throughput comes from the number of searches in flight (64 per SM).  The pattern is
// read from the current picture and the probes from the reference picture through L1/L2 — a search touches a few
// hundred scattered blocks of a window that 16 neighbouring CUs share.  Largest CUs are scheduled first.
#include "me_tz.cuh"

namespace vtmme {

namespace {

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

  TzCtx c;
  c.pat       = cur.origin + (ptrdiff_t) y * cur.stride + x;
  c.patStride = cur.stride;
  c.refAtPU   = ref.origin + (ptrdiff_t) y * ref.stride + x;
  c.refStride = ref.stride;
  c.w = c.h   = size;
  c.subShift  = (p.subShiftMode == 2 && size > 8 && size <= 64) ? 1 : 0;   // RdCost.cpp:310-316
  c.predQx    = pr.x;
  c.predQy    = pr.y;
  c.imvShift  = p.imvShift;
  c.lambda    = p.lambda;
  c.sm        = nullptr;

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
  const unsigned long long key = tz_search<1>(c, t);
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
