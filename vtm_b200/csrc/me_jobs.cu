// Per-call motion search (vtmme_search): one job = the integer full search of one
// InterSearch::xMotionEstimation call (xPatternSearch, EncoderLib/InterSearch.cpp:3566-3608) followed by
// the xPatternSearchFracDIF body (:4296-4338).  Any w,h in {4..128}, any window, row sub-sampling
// (DistParam::subShift) and signed patterns (bi-pred 2*org - otherPred, :3317-3328).
//
//   me_job_sad_kernel     CTA = (job, 32x32 sub-block of the pattern, slice of window rows): SADs with signed
//                         VABSDIFF, 8 displacements x 1 row per thread step.  Single-sub-block jobs keep their
//                         argmin directly; larger patterns write one SAD surface per sub-block.
//   me_job_reduce_kernel  sums the sub-block surfaces of a large pattern, adds lambda*bits, argmin.
//   me_job_tz_kernel      the integer search as xTZSearch (FastSearch=1/3, me_tz.cuh) instead of the full search: one CTA
//                         per job, leaves the same (cost, position) key for the refinement stage.
//   me_job_frac_kernel    fractional refinement (me_frac.cuh) or integer / 4-pel AMVR refinement (me_intrefine.cuh,
//                         fracMode 2 = xPatternSearchIntRefine) and result write-out.
#include "me_frac.cuh"
#include "me_intrefine.cuh"
#include "me_tz.cuh"
#include "me_kernels.h"

namespace vtmme {

namespace {

constexpr int kJobThreads  = 256;
constexpr int kJobBandRows = 32;   // maximum; small calls use 16 or 8 so that one job still spreads over many SMs

struct JobSmemHdr
{
  unsigned long long best;
  int                pad[2];
};
constexpr int kJobOffOrg = 16;                   // int32 [32][32]
constexpr int kJobOffRef = kJobOffOrg + 4096;    // int16 [rows][refStride]

__device__ __forceinline__ void job_consider(const DevJob& j, unsigned long long* best, int dx, int dy, uint32_t sad)
{
  if (dx < j.l || dx > j.r || dy < j.t || dy > j.b) return;
  const uint32_t cost = sad + mv_cost(j.lambda, mv_bits_q(dx * 4, dy * 4, j.predQx, j.predQy, j.imvShift));
  atomicMin(best, make_key(cost, dx, dy));
}

__device__ __forceinline__ int regions_of(const DevJob& j) { return ((j.w + 31) >> 5) * ((j.h + 31) >> 5); }

__global__ void __launch_bounds__(kJobThreads, 2) me_job_sad_kernel(const DevJob* __restrict__ jobs,
                                                                    unsigned long long* __restrict__ keys,
                                                                    uint32_t* __restrict__ surf,
                                                                    const long long* __restrict__ surfOff, int nSplit, int bandRows)
{
  extern __shared__ __align__(16) unsigned char smem[];
  JobSmemHdr* hdr   = reinterpret_cast<JobSmemHdr*>(smem);
  int32_t*    s_org = reinterpret_cast<int32_t*>(smem + kJobOffOrg);
  int16_t*    s_ref = reinterpret_cast<int16_t*>(smem + kJobOffRef);

  const DevJob j      = jobs[blockIdx.y];
  const int    nReg   = regions_of(j);
  const int    region = blockIdx.x / nSplit, split = blockIdx.x % nSplit;
  if (region >= nReg) return;
  const int tid = threadIdx.x;
  const int regX = (j.w + 31) >> 5;
  const int rx0 = (region % regX) * 32, ry0 = (region / regX) * 32;
  const int rw = min(32, j.w), rh = min(32, j.h);
  const int wl8 = j.l & ~7;
  const int ngx = (j.r - wl8 + 8) >> 3, nrows = j.b - j.t + 1;
  const int refStride = ngx * 8 + 40;
  const int step = 1 << j.subShift;
  const bool direct = nReg == 1;

  if (tid == 0) hdr->best = ~0ull;
  for (int i = tid; i < rw * rh; i += kJobThreads)
  {
    const int y = i / rw, x = i - y * rw;
    s_org[y * 32 + x] = (int32_t) j.org[(size_t) (ry0 + y) * j.orgStride + rx0 + x];
  }
  uint32_t* mySurf = direct ? nullptr : surf + surfOff[blockIdx.y] + (long long) region * nrows * (ngx * 8);

  for (int band0 = split * bandRows; band0 < nrows; band0 += nSplit * bandRows)
  {
    const int bh = min(bandRows, nrows - band0);
    __syncthreads();
    {
      // rows [j.t+band0, +bh+rh-1), cols [wl8, wl8 + ngx*8 + rw + 7] relative to the sub-block position
      const int cols = ngx * 8 + rw + 8;
      const int16_t* src = j.refAtPU + (ptrdiff_t) (ry0 + j.t + band0) * j.refStride + (rx0 + wl8);
      for (int i = tid; i < (bh + rh - 1) * cols; i += kJobThreads)
      {
        const int r = i / cols, c = i - r * cols;
        s_ref[r * refStride + c] = src[(ptrdiff_t) r * j.refStride + c];
      }
    }
    __syncthreads();
    const int ntiles = ngx * bh;
    for (int t = tid; t < ntiles; t += kJobThreads)
    {
      const int      dyi = t / ngx, gx = t - dyi * ngx;
      const int      dy = j.t + band0 + dyi, dx0 = wl8 + gx * 8;
      const int16_t* refTile = s_ref + dyi * refStride + gx * 8;
      uint32_t       a[8];
#pragma unroll
      for (int k = 0; k < 8; k++) a[k] = 0;
      for (int r = 0; r < rh; r += step)
      {
        if (rw >= 8)
        {
          for (int g = 0; g < rw; g += 8)
          {
            const int4  o0 = *reinterpret_cast<const int4*>(s_org + r * 32 + g);
            const int4  o1 = *reinterpret_cast<const int4*>(s_org + r * 32 + g + 4);
            const uint4 w0 = *reinterpret_cast<const uint4*>(refTile + r * refStride + g);
            const uint4 w1 = *reinterpret_cast<const uint4*>(refTile + r * refStride + g + 8);
            const int   o[8]   = { o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w };
            const int   px[16] = { (int) (short) (w0.x & 0xffffu), (int) w0.x >> 16, (int) (short) (w0.y & 0xffffu), (int) w0.y >> 16,
                                   (int) (short) (w0.z & 0xffffu), (int) w0.z >> 16, (int) (short) (w0.w & 0xffffu), (int) w0.w >> 16,
                                   (int) (short) (w1.x & 0xffffu), (int) w1.x >> 16, (int) (short) (w1.y & 0xffffu), (int) w1.y >> 16,
                                   (int) (short) (w1.z & 0xffffu), (int) w1.z >> 16, (int) (short) (w1.w & 0xffffu), (int) w1.w >> 16 };
#pragma unroll
            for (int k = 0; k < 8; k++)
#pragma unroll
              for (int i = 0; i < 8; i++) a[k] = __sad(o[i], px[i + k], a[k]);
          }
        }
        else
        {
          const int4  o0 = *reinterpret_cast<const int4*>(s_org + r * 32);
          const uint4 w0 = *reinterpret_cast<const uint4*>(refTile + r * refStride);
          const uint2 w1 = *reinterpret_cast<const uint2*>(refTile + r * refStride + 8);
          const int   o[4]   = { o0.x, o0.y, o0.z, o0.w };
          const int   px[12] = { (int) (short) (w0.x & 0xffffu), (int) w0.x >> 16, (int) (short) (w0.y & 0xffffu), (int) w0.y >> 16,
                                 (int) (short) (w0.z & 0xffffu), (int) w0.z >> 16, (int) (short) (w0.w & 0xffffu), (int) w0.w >> 16,
                                 (int) (short) (w1.x & 0xffffu), (int) w1.x >> 16, (int) (short) (w1.y & 0xffffu), (int) w1.y >> 16 };
#pragma unroll
          for (int k = 0; k < 8; k++)
#pragma unroll
            for (int i = 0; i < 4; i++) a[k] = __sad(o[i], px[i + k], a[k]);
        }
      }
      if (direct)
      {
        const uint32_t thr = (uint32_t) (*reinterpret_cast<volatile unsigned long long*>(&hdr->best) >> 32);
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
          const uint32_t sad = a[k] << j.subShift;
          if (sad <= thr) job_consider(j, &hdr->best, dx0 + k, dy, sad);
        }
      }
      else
      {
        uint4* dst = reinterpret_cast<uint4*>(mySurf + (size_t) (band0 + dyi) * (ngx * 8) + gx * 8);
        dst[0]     = make_uint4(a[0], a[1], a[2], a[3]);
        dst[1]     = make_uint4(a[4], a[5], a[6], a[7]);
      }
    }
  }
  __syncthreads();
  if (direct && tid == 0 && hdr->best != ~0ull) atomicMin(keys + blockIdx.y, hdr->best);
}

__global__ void __launch_bounds__(256) me_job_reduce_kernel(const DevJob* __restrict__ jobs,
                                                            unsigned long long* __restrict__ keys,
                                                            const uint32_t* __restrict__ surf,
                                                            const long long* __restrict__ surfOff)
{
  const DevJob j    = jobs[blockIdx.y];
  const int    nReg = regions_of(j);
  if (nReg == 1) return;
  const int wl8 = j.l & ~7;
  const int ngx = (j.r - wl8 + 8) >> 3, nrows = j.b - j.t + 1;
  const int ww = j.r - j.l + 1;
  const uint32_t* base = surf + surfOff[blockIdx.y];
  const size_t    regStride = (size_t) nrows * ngx * 8;
  unsigned long long best = ~0ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ww * nrows; i += gridDim.x * blockDim.x)
  {
    const int dyi = i / ww, dxi = i - dyi * ww;
    const int dx = j.l + dxi, dy = j.t + dyi;
    const size_t off = (size_t) dyi * (ngx * 8) + (dx - wl8);
    uint32_t     s = 0;
    for (int r = 0; r < nReg; r++) s += base[r * regStride + off];
    s <<= j.subShift;
    if (s <= key_cost(best))
    {
      const uint32_t cost = s + mv_cost(j.lambda, mv_bits_q(dx * 4, dy * 4, j.predQx, j.predQy, j.imvShift));
      const unsigned long long k = make_key(cost, dx, dy);
      if (k < best) best = k;
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1)
  {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, m);
    best                       = o < best ? o : best;
  }
  if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(keys + blockIdx.y, best);
}

// TZ search jobs: one CTA per job, pattern staged in shared memory (row stride w), reference read through L1/L2
__global__ void __launch_bounds__(kTzThreads) me_job_tz_kernel(const DevJob* __restrict__ jobs, const DevTz* __restrict__ tz,
                                                               unsigned long long* __restrict__ keys)
{
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ TzSmem sm;
  int16_t*     s_pat = reinterpret_cast<int16_t*>(smem);
  const DevJob j     = jobs[blockIdx.x];
  const DevTz  t     = tz[blockIdx.x];
  for (int i = threadIdx.x; i < j.w * j.h; i += kTzThreads)
  {
    const int y = i / j.w, x = i - y * j.w;
    s_pat[i]    = j.org[(size_t) y * j.orgStride + x];
  }
  __syncthreads();
  TzCtx c;
  c.pat       = s_pat;
  c.patStride = j.w;
  c.refAtPU   = j.refAtPU;
  c.refStride = j.refStride;
  c.w         = j.w;
  c.h         = j.h;
  c.subShift  = j.subShift;
  c.predQx    = j.predQx;
  c.predQy    = j.predQy;
  c.imvShift  = j.imvShift;
  c.lambda    = j.lambda;
  c.sm        = &sm;
  c.staged    = t.staged;
  const unsigned long long key = t.selective ? tz_search_selective<TzEvalWarps<kTzThreads / 32>>(c, t)
                                             : tz_search<TzEvalWarps<kTzThreads / 32>>(c, t);
  if (threadIdx.x == 0) keys[blockIdx.x] = key;
}

__device__ __forceinline__ FracJob make_frac_job(const DevJob& j, int dx, int dy)
{
  FracJob f;
  f.org        = j.org;
  f.orgStride  = j.orgStride;
  f.refAtMv    = j.refAtPU + (ptrdiff_t) dy * j.refStride + dx;
  f.refStride  = j.refStride;
  f.w          = j.w;
  f.h          = j.h;
  f.mvX        = dx;
  f.mvY        = dy;
  f.predQx     = j.predQx;
  f.predQy     = j.predQy;
  f.bitDepth   = j.bitDepth;
  f.useHad     = j.useHad;
  f.useAltHpel = j.useAltHpel;
  f.imvShift   = j.imvShift;
  f.lambda     = j.lambda;
  return f;
}

// Fractional refinement of the per-call jobs.  Patterns of one 32x32 chunk (the common case) run the whole
// xPatternSearchFracDIF body in one CTA (PASS 0).  Larger patterns are spread over one CTA per chunk: PASS 0 sums the
// half-pel candidates of each chunk into acc[job][0..8], PASS 1 takes the half-pel decision from those sums and adds the
// quarter-pel candidates to acc[job][9..17], me_job_frac_finish_kernel decides and writes the result.
template <int PASS>
__global__ void __launch_bounds__(kFracThreads) me_job_frac_kernel(const DevJob* __restrict__ jobs,
                                                                   unsigned long long* __restrict__ keys,
                                                                   DevJobResult* __restrict__ results,
                                                                   uint32_t* __restrict__ acc)
{
  __shared__ FracSmem   sm;
  __shared__ IntRefSmem irs;
  const int    job = blockIdx.y, chunk = blockIdx.x;
  const DevJob j   = jobs[job];
  const int    nChunks = frac_num_chunks(j.w, j.h);
  if (chunk >= nChunks) return;
  const unsigned long long key = keys[job];
  const int                dx = key_dx(key), dy = key_dy(key);
  const bool               single = nChunks == 1 || !j.fracMode;
  if (single)
  {
    if (PASS != 0 || chunk != 0) return;
    __syncthreads();
    if (threadIdx.x == 0) keys[job] = ~0ull;   // leave the slot ready for the next call (no memset per call)
    const uint32_t intCost = key_cost(key);
    DevJobResult   res;
    res.mvX    = dx;
    res.mvY    = dy;
    res.intSad = intCost - mv_cost(j.lambda, mv_bits_q(dx * 4, dy * 4, j.predQx, j.predQy, j.imvShift));
    res.halfX = res.halfY = res.qterX = res.qterY = 0;
    res.fracCost = res.intSad;
    res.amvrMvX = res.amvrMvY = res.mvpIdx = 0;
    res.bits = 0;
    res.cost = 0;
    if (j.fracMode == 2)
    {
      intrefine_accumulate<kFracThreads>(irs, j, j.org, j.orgStride, dx, dy, 0, 1);
      if (threadIdx.x == 0) intrefine_decide(j, dx, dy, irs.acc, res);
    }
    else if (j.fracMode)
    {
      const FracJob f = make_frac_job(j, dx, dy);
      if (j.imvShift > 1)
      {
        // xPatternSearchFracDIF :4311-4317 — integer AMVR: one SATD/SAD at the integer MV, rate at cost scale 2
        frac_stage(sm, f, 0, 0, 2, false);
        res.fracCost = sm.centre + mv_cost(j.lambda, mv_bits_q(dx * 4, dy * 4, j.predQx, j.predQy, j.imvShift));
      }
      else
      {
        const FracOut o = frac_refine_cta(sm, f);
        res.halfX    = o.halfX;
        res.halfY    = o.halfY;
        res.qterX    = o.qterX;
        res.qterY    = o.qterY;
        res.fracCost = o.cost;
      }
    }
    if (threadIdx.x == 0)
    {
      results[job] = res;   // mapped pinned host memory
      __threadfence_system();
    }
    return;
  }
  // ---- one chunk of a large pattern
  const FracJob f = make_frac_job(j, dx, dy);
  uint32_t*     a = acc + (size_t) job * 18;
  if (j.fracMode == 2)
  {
    if (PASS == 0)
    {
      intrefine_accumulate<kFracThreads>(irs, j, j.org, j.orgStride, dx, dy, chunk, nChunks);
      if (threadIdx.x < kIntRefProbes && irs.acc[threadIdx.x] != 0xffffffffu && irs.acc[threadIdx.x])
        atomicAdd(&a[threadIdx.x], irs.acc[threadIdx.x]);
    }
    return;
  }
  if (PASS == 0)
  {
    frac_stage_sums(sm, f, 0, 0, 2, j.useAltHpel != 0, chunk);
    if (threadIdx.x < 9) atomicAdd(&a[threadIdx.x], sm.acc[threadIdx.x]);
  }
  else if (j.imvShift == 0)
  {
    if (threadIdx.x == 0)
    {
      int      dir;
      uint32_t cost;
      frac_pick(a, f, 0, 0, 2, dir, cost);
      sm.best = dir;
    }
    __syncthreads();
    const int hx = c_refineH[sm.best][0], hy = c_refineH[sm.best][1];
    __syncthreads();
    frac_stage_sums(sm, f, hx * 2, hy * 2, 1, false, chunk);
    if (threadIdx.x < 9) atomicAdd(&a[9 + threadIdx.x], sm.acc[threadIdx.x]);
  }
}

// Decisions and result of the large patterns; resets their key and sums for the next call.
__global__ void __launch_bounds__(128) me_job_frac_finish_kernel(const DevJob* __restrict__ jobs,
                                                                 unsigned long long* __restrict__ keys,
                                                                 DevJobResult* __restrict__ results, uint32_t* __restrict__ acc,
                                                                 int n)
{
  const int job = blockIdx.x * blockDim.x + threadIdx.x;
  if (job >= n) return;
  const DevJob j = jobs[job];
  if (frac_num_chunks(j.w, j.h) == 1 || !j.fracMode) return;
  const unsigned long long key = keys[job];
  keys[job] = ~0ull;
  const int      dx = key_dx(key), dy = key_dy(key);
  const FracJob  f  = make_frac_job(j, dx, dy);
  uint32_t*      a  = acc + (size_t) job * 18;
  DevJobResult   res;
  res.mvX    = dx;
  res.mvY    = dy;
  res.intSad = key_cost(key) - mv_cost(j.lambda, mv_bits_q(dx * 4, dy * 4, j.predQx, j.predQy, j.imvShift));
  res.halfX = res.halfY = res.qterX = res.qterY = 0;
  res.amvrMvX = res.amvrMvY = res.mvpIdx = 0;
  res.bits = 0;
  res.cost = 0;
  if (j.fracMode == 2)
  {
    res.fracCost = res.intSad;
    intrefine_decide(j, dx, dy, a, res);
  }
  else if (j.imvShift > 1)
    res.fracCost = a[0] + mv_cost(j.lambda, mv_bits_q(dx * 4, dy * 4, j.predQx, j.predQy, j.imvShift));
  else
  {
    int      dir;
    uint32_t cost;
    frac_pick(a, f, 0, 0, 2, dir, cost);
    res.halfX = c_refineH[dir][0];
    res.halfY = c_refineH[dir][1];
    if (j.imvShift == 0)
    {
      frac_pick(a + 9, f, res.halfX * 2, res.halfY * 2, 1, dir, cost);
      res.qterX = c_refineQ[dir][0];
      res.qterY = c_refineQ[dir][1];
    }
    res.fracCost = cost;
  }
  for (int i = 0; i < 18; i++) a[i] = 0;
  results[job] = res;   // mapped pinned host memory
  __threadfence_system();
}


// Refinement of a pattern of at most 32x32 samples by one CTA of kFracThreads threads, the pattern already in shared
// memory (row stride patStride): the xPatternSearchFracDIF body (fracMode 1) or xPatternSearchIntRefine (fracMode 2)
// at the integer MV held by `key`, and the result record (`out`: mapped pinned host memory).
__device__ __forceinline__ void job_refine_single(const DevJob& j, unsigned long long key, const int16_t* s_pat, int patStride,
                                                  FracSmem& fsm, IntRefSmem& irs, DevJobResult* out, unsigned int* done,
                                                  unsigned int seq)
{
  const int    dx = key_dx(key), dy = key_dy(key);
  DevJobResult res;
  res.mvX    = dx;
  res.mvY    = dy;
  res.intSad = key_cost(key) - mv_cost(j.lambda, mv_bits_q(dx * 4, dy * 4, j.predQx, j.predQy, j.imvShift));
  res.halfX = res.halfY = res.qterX = res.qterY = 0;
  res.fracCost = res.intSad;
  res.amvrMvX = res.amvrMvY = res.mvpIdx = 0;
  res.bits = 0;
  res.cost = 0;
  if (j.fracMode == 2)
  {
    intrefine_accumulate<kFracThreads>(irs, j, s_pat, patStride, dx, dy, 0, 1);
    if (threadIdx.x == 0) intrefine_decide(j, dx, dy, irs.acc, res);
  }
  else if (j.fracMode)
  {
    FracJob f   = make_frac_job(j, dx, dy);
    f.org       = s_pat;
    f.orgStride = patStride;
    if (j.imvShift > 1)
    {
      frac_stage(fsm, f, 0, 0, 2, false);
      res.fracCost = fsm.centre + mv_cost(j.lambda, mv_bits_q(dx * 4, dy * 4, j.predQx, j.predQy, j.imvShift));
    }
    else
    {
      const FracOut o = frac_refine_cta(fsm, f);
      res.halfX    = o.halfX;
      res.halfY    = o.halfY;
      res.qterX    = o.qterX;
      res.qterY    = o.qterY;
      res.fracCost = o.cost;
    }
  }
  if (threadIdx.x == 0)
  {
    *out = res;
    __threadfence_system();
    // the host polls this word instead of waiting for the stream: the result above is visible before it
    if (done) *reinterpret_cast<volatile unsigned int*>(done) = seq;
  }
}

// TZ search of one small job with its refinement, one launch (the in-loop encoder's FastSearch=1 call)
__global__ void __launch_bounds__(kTzThreads) me_job_tz_fused_kernel(const DevJob* __restrict__ jobs, const DevTz* __restrict__ tz,
                                                                     DevJobResult* __restrict__ result, unsigned int* done,
                                                                     unsigned int seq)
{
  static_assert(kTzThreads == kFracThreads, "the refinement code is written for CTAs of kFracThreads threads");
  __shared__ TzSmem                sm;
  __shared__ FracSmem              fsm;
  __shared__ IntRefSmem            irs;
  __shared__ __align__(16) int16_t s_pat[32 * 32];
  const DevJob j = jobs[0];
  const DevTz  t = tz[0];
  for (int i = threadIdx.x; i < j.w * j.h; i += kTzThreads)
  {
    const int y = i / j.w, x = i - y * j.w;
    s_pat[i]    = j.org[(size_t) y * j.orgStride + x];
  }
  __syncthreads();
  TzCtx c;
  c.pat       = s_pat;
  c.patStride = j.w;
  c.refAtPU   = j.refAtPU;
  c.refStride = j.refStride;
  c.w         = j.w;
  c.h         = j.h;
  c.subShift  = j.subShift;
  c.predQx    = j.predQx;
  c.predQy    = j.predQy;
  c.imvShift  = j.imvShift;
  c.lambda    = j.lambda;
  c.sm        = &sm;
  c.staged    = t.staged;
  const unsigned long long key = t.selective ? tz_search_selective<TzEvalWarps<kTzThreads / 32>>(c, t)
                                             : tz_search<TzEvalWarps<kTzThreads / 32>>(c, t);
  __syncthreads();
  job_refine_single(j, key, s_pat, j.w, fsm, irs, result, done, seq);
}

// ---- single small job, one launch (the in-loop encoder's call: one PU, pattern <= 32x32) ----------------------------
// Job descriptor and pattern travel as kernel parameters (no upload, no dependent global reads at kernel start); every
// CTA searches one band of window rows, the last CTA to finish (ticket) runs the fractional refinement and writes the
// result to mapped pinned memory.
constexpr int kFusedThreads = kFracThreads;   // 128: the refinement code is written for CTAs of kFracThreads threads

__global__ void __launch_bounds__(kFusedThreads) me_job_fused_kernel(const __grid_constant__ FusedJobArgs a)
{
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ FracSmem   fsm;
  __shared__ IntRefSmem irs;
  __shared__ int16_t    s_pat[32 * 32];
  __shared__ int      s_last;
  JobSmemHdr* hdr   = reinterpret_cast<JobSmemHdr*>(smem);
  int32_t*    s_org = reinterpret_cast<int32_t*>(smem + kJobOffOrg);
  int16_t*    s_ref = reinterpret_cast<int16_t*>(smem + kJobOffRef);

  const DevJob& j   = a.job;
  const int     tid = threadIdx.x;
  const int rw = j.w, rh = j.h;
  const int wl8 = j.l & ~7;
  const int ngx = (j.r - wl8 + 8) >> 3, nrows = j.b - j.t + 1;
  const int refStride = ngx * 8 + 40;
  const int step = 1 << j.subShift;

  if (tid == 0) hdr->best = ~0ull;
  for (int i = tid; i < rw * rh; i += kFusedThreads)
  {
    const int     y = i / rw, x = i - y * rw;
    const int16_t v = a.inlinePattern ? a.pattern[i] : j.org[(size_t) y * j.orgStride + x];
    s_org[y * 32 + x] = (int32_t) v;
    s_pat[i]          = v;
  }
  for (int band0 = blockIdx.x * a.bandRows; band0 < nrows; band0 += gridDim.x * a.bandRows)
  {
    const int bh = min(a.bandRows, nrows - band0);
    __syncthreads();
    {
      const int cols = ngx * 8 + rw + 8;
      const int16_t* src = j.refAtPU + (ptrdiff_t) (j.t + band0) * j.refStride + wl8;
      for (int i = tid; i < (bh + rh - 1) * cols; i += kFusedThreads)
      {
        const int r = i / cols, c = i - r * cols;
        s_ref[r * refStride + c] = src[(ptrdiff_t) r * j.refStride + c];
      }
    }
    __syncthreads();
    const int ntiles = ngx * bh;
    for (int t = tid; t < ntiles; t += kFusedThreads)
    {
      const int      dyi = t / ngx, gx = t - dyi * ngx;
      const int      dy = j.t + band0 + dyi, dx0 = wl8 + gx * 8;
      const int16_t* refTile = s_ref + dyi * refStride + gx * 8;
      uint32_t       acc[8];
#pragma unroll
      for (int k = 0; k < 8; k++) acc[k] = 0;
      for (int r = 0; r < rh; r += step)
      {
        if (rw >= 8)
        {
          for (int g = 0; g < rw; g += 8)
          {
            const int4  o0 = *reinterpret_cast<const int4*>(s_org + r * 32 + g);
            const int4  o1 = *reinterpret_cast<const int4*>(s_org + r * 32 + g + 4);
            const uint4 w0 = *reinterpret_cast<const uint4*>(refTile + r * refStride + g);
            const uint4 w1 = *reinterpret_cast<const uint4*>(refTile + r * refStride + g + 8);
            const int   o[8]   = { o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w };
            const int   px[16] = { (int) (short) (w0.x & 0xffffu), (int) w0.x >> 16, (int) (short) (w0.y & 0xffffu), (int) w0.y >> 16,
                                   (int) (short) (w0.z & 0xffffu), (int) w0.z >> 16, (int) (short) (w0.w & 0xffffu), (int) w0.w >> 16,
                                   (int) (short) (w1.x & 0xffffu), (int) w1.x >> 16, (int) (short) (w1.y & 0xffffu), (int) w1.y >> 16,
                                   (int) (short) (w1.z & 0xffffu), (int) w1.z >> 16, (int) (short) (w1.w & 0xffffu), (int) w1.w >> 16 };
#pragma unroll
            for (int k = 0; k < 8; k++)
#pragma unroll
              for (int i = 0; i < 8; i++) acc[k] = __sad(o[i], px[i + k], acc[k]);
          }
        }
        else
        {
          const int4  o0 = *reinterpret_cast<const int4*>(s_org + r * 32);
          const uint4 w0 = *reinterpret_cast<const uint4*>(refTile + r * refStride);
          const uint2 w1 = *reinterpret_cast<const uint2*>(refTile + r * refStride + 8);
          const int   o[4]   = { o0.x, o0.y, o0.z, o0.w };
          const int   px[12] = { (int) (short) (w0.x & 0xffffu), (int) w0.x >> 16, (int) (short) (w0.y & 0xffffu), (int) w0.y >> 16,
                                 (int) (short) (w0.z & 0xffffu), (int) w0.z >> 16, (int) (short) (w0.w & 0xffffu), (int) w0.w >> 16,
                                 (int) (short) (w1.x & 0xffffu), (int) w1.x >> 16, (int) (short) (w1.y & 0xffffu), (int) w1.y >> 16 };
#pragma unroll
          for (int k = 0; k < 8; k++)
#pragma unroll
            for (int i = 0; i < 4; i++) acc[k] = __sad(o[i], px[i + k], acc[k]);
        }
      }
      const uint32_t thr = (uint32_t) (*reinterpret_cast<volatile unsigned long long*>(&hdr->best) >> 32);
#pragma unroll
      for (int k = 0; k < 8; k++)
      {
        const uint32_t sad = acc[k] << j.subShift;
        if (sad <= thr) job_consider(j, &hdr->best, dx0 + k, dy, sad);
      }
    }
  }
  __syncthreads();
  if (tid == 0)
  {
    if (hdr->best != ~0ull) atomicMin(a.key, hdr->best);
    __threadfence();
    s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;

  // ---- last CTA: the window is complete.  xPatternSearchFracDIF body and result.
  if (tid == 0)
  {
    *a.ticket = 0;                                                   // ready for the next call
    hdr->best = atomicExch(a.key, ~0ull);                            // coherent read + reset of the slot
  }
  __syncthreads();
  const unsigned long long key = hdr->best;
  const int                dx = key_dx(key), dy = key_dy(key);
  job_refine_single(j, key, s_pat, rw, fsm, irs, a.result, a.done, a.seq);
}

}   // namespace

// dJobs must be followed in the same buffer by nothing the kernels need; dKeys [n]; surfaces are sized and
// offset by the caller (vtmme_api.cu) and passed through dSurf / dSurfOff.
cudaError_t launch_job_search_impl(const DevJob* dJobs, unsigned long long* dKeys, DevJobResult* dResults, int n,
                                   int maxRegions, int nSplit, int bandRows, int maxGx, bool anyMulti, uint32_t* dSurf,
                                   const long long* dSurfOff, uint32_t* dFracAcc, int maxFracChunks, cudaStream_t st,
                                   int* launches, const DevTz* dTz, int maxPatternSamples, unsigned int* done,
                                   unsigned int seq, bool* fusedTz)
{
  cudaError_t e;
  if (fusedTz) *fusedTz = false;
  if (dTz && n == 1 && maxPatternSamples <= 32 * 32 && maxRegions == 1)
  {
    // one small TZ job: search and refinement in one launch
    if (fusedTz) *fusedTz = true;
    me_job_tz_fused_kernel<<<1, kTzThreads, 0, st>>>(dJobs, dTz, dResults, done, seq);
    *launches += 1;
    return cudaGetLastError();
  }
  if (dTz)
  {
    me_job_tz_kernel<<<n, kTzThreads, (size_t) maxPatternSamples * 2, st>>>(dJobs, dTz, dKeys);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    *launches += 1;
  }
  else
  {
    const size_t smem = (size_t) kJobOffRef + (size_t) (kJobBandRows + 31) * (maxGx * 8 + 40) * 2;
    static SmemOptIn optIn;
    if ((e = optIn.ensure(me_job_sad_kernel, smem)) != cudaSuccess) return e;
    dim3 grid(maxRegions * nSplit, n, 1);
    me_job_sad_kernel<<<grid, kJobThreads, smem, st>>>(dJobs, dKeys, dSurf, dSurfOff, nSplit, bandRows);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    *launches += 1;
  }
  if (anyMulti && !dTz)
  {
    dim3 g2(16, n, 1);
    me_job_reduce_kernel<<<g2, 256, 0, st>>>(dJobs, dKeys, dSurf, dSurfOff);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    *launches += 1;
  }
  dim3 gf(maxFracChunks, n, 1);
  me_job_frac_kernel<0><<<gf, kFracThreads, 0, st>>>(dJobs, dKeys, dResults, dFracAcc);
  *launches += 1;
  if (maxFracChunks > 1)
  {
    me_job_frac_kernel<1><<<gf, kFracThreads, 0, st>>>(dJobs, dKeys, dResults, dFracAcc);
    me_job_frac_finish_kernel<<<(n + 127) / 128, 128, 0, st>>>(dJobs, dKeys, dResults, dFracAcc, n);
    *launches += 2;
  }
  return cudaGetLastError();
}

size_t fused_job_smem_bytes(const FusedJobArgs& a)
{
  const int wl8 = a.job.l & ~7, ngx = (a.job.r - wl8 + 8) >> 3;
  return (size_t) kJobOffRef + (size_t) (a.bandRows + a.job.h - 1) * (ngx * 8 + 40) * 2;
}

cudaError_t launch_job_fused(const FusedJobArgs& a, int grid, cudaStream_t st)
{
  me_job_fused_kernel<<<grid, kFusedThreads, fused_job_smem_bytes(a), st>>>(a);
  return cudaGetLastError();
}

}   // namespace vtmme
