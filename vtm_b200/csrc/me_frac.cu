// Fractional refinement + result write-out of the batched frame path (one CTA per CU).
#include "me_frac.cuh"
#include "me_kernels.h"

#include "../../include/vtmme.h"

namespace vtmme {

namespace {

__global__ void __launch_bounds__(kFracThreads) me_frac_frame_kernel(FracFrameParams p)
{
  __shared__ FracSmem sm;
  const int cu = blockIdx.x, pair = blockIdx.y;
  const int nCU = p.g.off[5];
  int       level = 0;
#pragma unroll
  for (int l = 1; l < 5; l++)
    if (cu >= p.g.off[l]) level = l;
  const int size = 8 << level;
  const int li = cu - p.g.off[level];
  const int cx = li % p.g.nx[level], cy = li / p.g.nx[level];
  const int x = cx * size, y = cy * size;

  const unsigned long long key = p.keys[(size_t) pair * nCU + cu];
  short2 pr = make_short2(0, 0);
  if (p.predQ) pr = p.predQ[(size_t) pair * nCU + cu];
  const int      dx = key_dx(key), dy = key_dy(key);
  const uint32_t intCost = key_cost(key);
  const uint32_t intSad  = intCost - mv_cost(p.lambda, mv_bits_q(dx * 4, dy * 4, pr.x, pr.y, p.imvShift));

  vtmme_cu_result res;
  res.intX     = (int16_t) dx;
  res.intY     = (int16_t) dy;
  res.intSad   = intSad;
  res.mvQx     = (int16_t) (dx * 4);
  res.mvQy     = (int16_t) (dy * 4);
  res.fracCost = intCost;
  if (p.fracMode)
  {
    const DevPic cur = p.cur[pair], ref = p.ref[pair];
    FracJob      j;
    j.org        = cur.origin + (size_t) y * cur.stride + x;
    j.orgStride  = cur.stride;
    j.refAtMv    = ref.origin + (ptrdiff_t) (y + dy) * ref.stride + (x + dx);
    j.refStride  = ref.stride;
    j.w = j.h    = size;
    j.mvX        = dx;
    j.mvY        = dy;
    j.predQx     = pr.x;
    j.predQy     = pr.y;
    j.bitDepth   = p.bitDepth;
    j.useHad     = p.useHad;
    j.useAltHpel = 0;
    j.imvShift   = p.imvShift;
    j.lambda     = p.lambda;
    const FracOut o = frac_refine_cta(sm, j);
    res.mvQx     = (int16_t) (dx * 4 + o.halfX * 2 + o.qterX);
    res.mvQy     = (int16_t) (dy * 4 + o.halfY * 2 + o.qterY);
    res.fracCost = o.cost;
  }
  if (threadIdx.x == 0) reinterpret_cast<vtmme_cu_result*>(p.results)[(size_t) pair * nCU + cu] = res;
}

}   // namespace

cudaError_t launch_frac_frame(const FracFrameParams& p, int nPairs, cudaStream_t st)
{
  dim3 grid(p.g.off[5], nPairs, 1);
  me_frac_frame_kernel<<<grid, kFracThreads, 0, st>>>(p);
  return cudaGetLastError();
}

}   // namespace vtmme
