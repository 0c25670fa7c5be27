// Batched TZ search of the frame path with the search windows staged in shared memory (levels 8x8 .. 64x64).
//
// me_tz.cu reads every probe straight from the reference picture: a probe of an 8x8 CU is eight 16-byte rows in eight
// different cache lines, and the level is bound by L1 tag look-ups (one per distinct line and instruction).  A TZ search,
// however, only ever probes inside two windows known before it starts — xSetSearchRange around its start point and around
// the zero vector (InterSearch.cpp:3767-3772; the start point is the predictor, the zero vector is tested next, :3695-3712)
// — and neighbouring CUs share most of them.  So a CTA takes a small group of neighbouring CUs, stages the bounding box of
// their windows once (coalesced 16-byte loads, one 32-bit word per sample), and then every probe is shared-memory reads:
//   * lane per probe (a raster scan is up to 676 probes, a diamond up to 16): a lane walks the rows of its probe, one
//     LDS.32 + one VABSDIFF per sample, the pattern row comes as broadcast LDS.128;
//   * when a batch has fewer probes than the search has lanes, a probe is cut into up to 32 row chunks over consecutive
//     lanes (summed with shuffles), so that small batches of large CUs still fill the lanes;
//   * the key (cost, order in the batch) is reduced with shuffles, and through one shared word for CUs searched by
//     several warps.
// The sequential control flow is the shared tz_search<EV> of me_tz.cuh; only the evaluator differs, so the result is
// the reference's by the same argument (first strict minimum in probe order).  A probe outside the staged box (cannot
// happen for the two windows above; kept as a guard) is read from the picture.
#include "me_tz.cuh"

namespace vtmme {

namespace {

template <int SIZE, int SS, int NW>   // NW warps cooperate on one search
struct TzEvalSmem
{
  static constexpr int ROWS = SIZE >> SS;   // sampled rows
  static constexpr int T    = NW * 32;
  static constexpr int kOrgStride = SIZE + 4;   // pattern rows staggered by 16 bytes
  struct Ctx
  {
    const uint32_t* win;        // staged window, one sample per word: sample (dx + c, dy + r) of the CU at win[(dy + r) * stride + dx + c]
    const uint32_t* org;        // pattern, one sample per word, row stride kOrgStride
    int             stride;
    int             bl, br, bt, bb;   // displacements the staged box covers for this CU
    const int16_t*  refAtPU;    // the picture, for the guard path
    int             refStride;
    int             predQx, predQy, imvShift;
    double          lambda;
    unsigned long long* key;    // NW > 1: shared slot
    int             tid;        // thread index within the search (0 .. T-1)
  };

  // columns [col0, col0 + nCols) of all sampled rows of the probe at displacement (x, y).  The lanes that share a probe
  // split its COLUMNS: they read consecutive words of the same window row and of the same pattern row, which is free of
  // bank conflicts whatever the row stride is (rows of a probe would collide: the stride is a multiple of four words).
  static __device__ __forceinline__ uint32_t unit_sad(const Ctx& c, int x, int y, int col0, int nCols)
  {
    uint32_t acc = 0, a1 = 0, a2 = 0, a3 = 0;   // independent chains: a single accumulator is bound by the ALU latency
    if (x >= c.bl && x <= c.br && y >= c.bt && y <= c.bb)
    {
      const uint32_t* w = c.win + y * c.stride + x + col0;
      const uint32_t* o = c.org + col0;
      const int       ws = c.stride << SS, os = kOrgStride << SS;
      if (nCols == SIZE)   // a whole probe per lane (raster scans): the row is unrolled
      {
        for (int r = 0; r < ROWS; r++, w += ws, o += os)
#pragma unroll
          for (int q = 0; q < SIZE; q += 4)
          {
            const uint4 ov = *reinterpret_cast<const uint4*>(o + q);
            acc = __usad(ov.x, w[q], acc);
            a1  = __usad(ov.y, w[q + 1], a1);
            a2  = __usad(ov.z, w[q + 2], a2);
            a3  = __usad(ov.w, w[q + 3], a3);
          }
      }
      else if (nCols >= 4)
      {
        for (int r = 0; r < ROWS; r++, w += ws, o += os)
          for (int q = 0; q < nCols; q += 4)
          {
            const uint4 ov = *reinterpret_cast<const uint4*>(o + q);
            acc = __usad(ov.x, w[q], acc);
            a1  = __usad(ov.y, w[q + 1], a1);
            a2  = __usad(ov.z, w[q + 2], a2);
            a3  = __usad(ov.w, w[q + 3], a3);
          }
      }
      else if (nCols == 2)
      {
#pragma unroll 4
        for (int r = 0; r < ROWS; r++, w += ws, o += os)
        {
          const uint2 ov = *reinterpret_cast<const uint2*>(o);
          acc = __usad(ov.x, w[0], acc);
          a1  = __usad(ov.y, w[1], a1);
        }
      }
      else
      {
#pragma unroll 4
        for (int r = 0; r < ROWS; r++, w += ws, o += os) acc = __usad(o[0], w[0], acc);
      }
    }
    else
    {
      const int16_t*  g = c.refAtPU + (ptrdiff_t) y * c.refStride + x + col0;
      const uint32_t* o = c.org + col0;
      for (int r = 0; r < ROWS; r++, g += (ptrdiff_t) c.refStride << SS, o += kOrgStride << SS)
        for (int q = 0; q < nCols; q++) acc = __usad(o[q], (uint32_t) (uint16_t) g[q], acc);
    }
    return (acc + a1) + (a2 + a3);
  }

  template <class PointFn>
  static __device__ __forceinline__ int eval_plain(const Ctx& c, int n, uint32_t best, uint32_t& costOut, PointFn point)
  {
    return eval(c, n, best, costOut, point);
  }

  template <class PointFn>
  static __device__ __forceinline__ int eval(const Ctx& c, int n, uint32_t best, uint32_t& costOut, PointFn point)
  {
    // column chunks per probe: as many as keep the lanes busy, at most 32 (a probe's chunks stay inside one warp)
    int C = 1, shiftC = 0;
    while (C * 2 <= SIZE && n * C * 2 <= T && C < 32)
    {
      C *= 2;
      shiftC++;
    }
    const int colsPer = SIZE >> shiftC;
    unsigned long long k = ~0ull;
    for (int u0 = 0; u0 < n * C; u0 += T)
    {
      const int  u = u0 + c.tid, i = u >> shiftC, ch = u & (C - 1);
      const bool valid = i < n;
      int        x, y;
      point(valid ? i : 0, x, y);
      uint32_t sad = valid ? unit_sad(c, x, y, ch * colsPer, colsPer) : 0u;
      for (int m = 1; m < C; m <<= 1) sad += __shfl_xor_sync(0xffffffffu, sad, m);
      if (valid && ch == 0)
      {
        const unsigned long long ki = ((unsigned long long) tz_cost(c, x, y, sad << SS) << 32) | (uint32_t) i;
        k = ki < k ? ki : k;
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1)
    {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, m);
      k                          = o < k ? o : k;
    }
    if (NW > 1)
    {
      __syncthreads();   // the previous batch's key has been read by everyone
      if (c.tid == 0) *c.key = ~0ull;
      __syncthreads();
      if ((c.tid & 31) == 0 && k != ~0ull) atomicMin(c.key, k);
      __syncthreads();
      k = *c.key;
    }
    costOut = (uint32_t) (k >> 32);
    return (k != ~0ull && costOut < best) ? (int) (uint32_t) k : -1;
  }
};

// GX x GY neighbouring CUs per CTA, NW warps per CU
template <int SIZE, int SS, int GX, int GY, int NW>
__global__ void __launch_bounds__(GX * GY * NW * 32) me_tz_frame_smem_kernel(TzFrameParams p, int level, int capWords)
{
  constexpr int NCU = GX * GY, THREADS = NCU * NW * 32;
  using EV = TzEvalSmem<SIZE, SS, NW>;
  extern __shared__ __align__(16) uint32_t s_mem[];
  __shared__ int                s_box[NCU][4];   // per CU: union of both windows (l, r, t, b) in displacements
  __shared__ int                s_boxP[NCU][4];  // per CU: the window around the start point only
  __shared__ int                s_abs[5];        // box in picture coordinates (x0, x1, y0, y1) and its row stride
  __shared__ unsigned long long s_key;
  constexpr int OS = SIZE + 4;                   // pattern row stride (EV::kOrgStride)
  uint32_t* s_org = s_mem;                       // [NCU][SIZE][OS]
  uint32_t* s_win = s_mem + NCU * SIZE * OS;     // the staged box

  const int tid = threadIdx.x, pair = blockIdx.y;
  const int nx = p.g.nx[level], ny = p.g.ny[level], nCU = p.g.off[5];
  const int groupsX = (nx + GX - 1) / GX;
  const int gx0 = (blockIdx.x % groupsX) * GX, gy0 = (blockIdx.x / groupsX) * GY;
  const DevPic cur = p.cur[pair], ref = p.ref[pair];

  // 1. the two windows of every CU of the group
  if (tid < NCU)
  {
    const int cx = gx0 + tid % GX, cy = gy0 + tid / GX;
    int l = 1 << 20, r = -(1 << 20), t = 1 << 20, b = -(1 << 20);
    int lp = l, rp = r, tp = t, bp = b;
    if (cx < nx && cy < ny)
    {
      short2 pr = make_short2(0, 0);
      if (p.predQ) pr = p.predQ[(size_t) pair * nCU + p.g.off[level] + cy * nx + cx];
      DevTz tt;
      tt.posX = cx * SIZE;
      tt.posY = cy * SIZE;
      tt.picW = p.g.picW;
      tt.picH = p.g.picH;
      tt.maxCuW = tt.maxCuH = p.ctu;
      int sx = pr.x * 4, sy = pr.y * 4;
      tz_to_int(tt, sx, sy);
      const Window w1 = search_window(sx * 4, sy * 4, tt.posX, tt.posY, p.g.picW, p.g.picH, p.ctu, p.sr);
      const Window w0 = search_window(0, 0, tt.posX, tt.posY, p.g.picW, p.g.picH, p.ctu, p.sr);
      lp = min(w1.l, sx);
      rp = max(w1.r, sx);
      tp = min(w1.t, sy);
      bp = max(w1.b, sy);
      l = min(min(lp, w0.l), 0);
      r = max(max(rp, w0.r), 0);
      t = min(min(tp, w0.t), 0);
      b = max(max(bp, w0.b), 0);
    }
    s_box[tid][0] = l;
    s_box[tid][1] = r;
    s_box[tid][2] = t;
    s_box[tid][3] = b;
    s_boxP[tid][0] = lp;
    s_boxP[tid][1] = rp;
    s_boxP[tid][2] = tp;
    s_boxP[tid][3] = bp;
  }
  __syncthreads();
  if (tid == 0)
  {
    // first choice: both windows of every CU; if that does not fit (predictors far from the zero vector), the windows
    // around the start points only — probes around the zero vector then take the guard path
    int stride = 0, x0 = 0, x1 = -1, y0 = 0, y1 = -1;
    for (int choice = 0; choice < 2 && x1 < x0; choice++)
    {
      x0 = 1 << 20, x1 = -(1 << 20), y0 = 1 << 20, y1 = -(1 << 20);
      for (int k = 0; k < NCU; k++)
      {
        const int* bx = choice == 0 ? s_box[k] : s_boxP[k];
        if (bx[1] < bx[0]) continue;
        const int px = (gx0 + k % GX) * SIZE, py = (gy0 + k / GX) * SIZE;
        x0 = min(x0, px + bx[0]);
        x1 = max(x1, px + bx[1] + SIZE - 1);
        y0 = min(y0, py + bx[2]);
        y1 = max(y1, py + bx[3] + SIZE - 1);
      }
      x0 &= ~7;                                 // 16-byte aligned rows in the picture
      stride = ((x1 - x0 + 1 + 7) & ~7) + 4;    // + 4 words: consecutive rows start in different bank groups
      if (x1 < x0 || (long long) stride * (y1 - y0 + 1) > capWords)
      {
        x1 = x0 - 1;                            // does not fit
        stride = 0;
      }
    }
    s_abs[0] = x0;
    s_abs[1] = x1;
    s_abs[2] = y0;
    s_abs[3] = y1;
    s_abs[4] = stride;
  }
  __syncthreads();
  const int bx0 = s_abs[0], bx1 = s_abs[1], by0 = s_abs[2], by1 = s_abs[3], wstride = s_abs[4];
  // 2. stage the box (8 samples per unit) and the patterns
  if (bx1 >= bx0)
  {
    const int upr = (bx1 - bx0 + 8) >> 3, rows = by1 - by0 + 1;
    for (int i = tid; i < upr * rows; i += THREADS)
    {
      const int   r = i / upr, u = i - r * upr;
      const uint4 v = *reinterpret_cast<const uint4*>(ref.origin + (ptrdiff_t) (by0 + r) * ref.stride + bx0 + 8 * u);
      uint32_t*   d = s_win + r * wstride + 8 * u;
      *reinterpret_cast<uint4*>(d)     = make_uint4(v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16);
      *reinterpret_cast<uint4*>(d + 4) = make_uint4(v.z & 0xffffu, v.z >> 16, v.w & 0xffffu, v.w >> 16);
    }
  }
  for (int i = tid; i < NCU * SIZE * SIZE / 8; i += THREADS)
  {
    const int k = i / (SIZE * SIZE / 8), rem = i - k * (SIZE * SIZE / 8), r = rem / (SIZE / 8), u = rem - r * (SIZE / 8);
    const int cx = gx0 + k % GX, cy = gy0 + k / GX;
    if (cx >= nx || cy >= ny) continue;
    const uint4 v = *reinterpret_cast<const uint4*>(cur.origin + (ptrdiff_t) (cy * SIZE + r) * cur.stride + cx * SIZE + 8 * u);
    uint32_t*   d = s_org + k * SIZE * OS + r * OS + 8 * u;
    *reinterpret_cast<uint4*>(d)     = make_uint4(v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16);
    *reinterpret_cast<uint4*>(d + 4) = make_uint4(v.z & 0xffffu, v.z >> 16, v.w & 0xffffu, v.w >> 16);
  }
  __syncthreads();

  // 3. one search per NW warps
  const int k = tid / (NW * 32);
  const int cx = gx0 + k % GX, cy = gy0 + k / GX;
  if (cx >= nx || cy >= ny) return;   // (NCU > 1 implies NW == 1: no block-wide barrier below)
  const int x = cx * SIZE, y = cy * SIZE, cu = p.g.off[level] + cy * nx + cx;
  short2    pr = make_short2(0, 0);
  if (p.predQ) pr = p.predQ[(size_t) pair * nCU + cu];
  typename EV::Ctx c;
  c.stride    = wstride;
  c.win       = s_win + (y - by0) * wstride + (x - bx0);
  c.org       = s_org + k * SIZE * OS;
  c.bl        = bx1 >= bx0 ? bx0 - x : 1;
  c.br        = bx1 >= bx0 ? bx1 - (SIZE - 1) - x : 0;
  c.bt        = by0 - y;
  c.bb        = by1 - (SIZE - 1) - y;
  c.refAtPU   = ref.origin + (ptrdiff_t) y * ref.stride + x;
  c.refStride = ref.stride;
  c.predQx    = pr.x;
  c.predQy    = pr.y;
  c.imvShift  = p.imvShift;
  c.lambda    = p.lambda;
  c.key       = &s_key;
  c.tid       = tid - k * NW * 32;

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
  t.selective = t.staged = 0;
  const unsigned long long key = tz_search<EV>(c, t);
  if (c.tid == 0) p.keys[(size_t) pair * nCU + cu] = key;
}

template <int SIZE, int SS, int GX, int GY, int NW>
cudaError_t launch_smem_level(const TzFrameParams& p, int level, int nPairs, int predSpread, cudaStream_t st, bool* done)
{
  static SmemOptIn optIn;
  *done = false;
  const int nx = p.g.nx[level], ny = p.g.ny[level];
  if (nx * ny == 0)
  {
    *done = true;
    return cudaSuccess;
  }
  // largest box of a group: both windows of every CU, predictors up to predSpread samples apart
  const long long bw = 2 * p.sr + 1 + predSpread + GX * SIZE + 7, bh = 2 * p.sr + 1 + predSpread + GY * SIZE - 1;
  const long long words = (((bw + 7) & ~7) + 4) * bh;
  const size_t smem = (size_t) (words + GX * GY * SIZE * (SIZE + 4)) * 4;
  if (smem > 200 * 1024) return cudaSuccess;   // does not fit: the caller takes the kernels of me_tz.cu for this level
  cudaError_t e = optIn.ensure(me_tz_frame_smem_kernel<SIZE, SS, GX, GY, NW>, smem);
  if (e != cudaSuccess) return e;
  dim3 grid(((nx + GX - 1) / GX) * ((ny + GY - 1) / GY), nPairs, 1);
  me_tz_frame_smem_kernel<SIZE, SS, GX, GY, NW><<<grid, GX * GY * NW * 32, smem, st>>>(p, level, (int) words);
  *done = true;
  return cudaGetLastError();
}

}   // namespace

// Level `level` (0 .. 3) of the frame TZ search with staged windows; *done tells whether the level was launched (false:
// the group's box does not fit shared memory for this search range / predictor spread).
cudaError_t launch_tz_frame_smem(const TzFrameParams& p, int level, int nPairs, int predSpread, cudaStream_t st, bool* done)
{
  const bool ss = p.subShiftMode == 2;
  switch (level)
  {
  case 0: return launch_smem_level<8, 0, 4, 2, 1>(p, 0, nPairs, predSpread, st, done);
  case 1: return ss ? launch_smem_level<16, 1, 2, 2, 1>(p, 1, nPairs, predSpread, st, done) : launch_smem_level<16, 0, 2, 2, 1>(p, 1, nPairs, predSpread, st, done);
  case 2: return ss ? launch_smem_level<32, 1, 1, 1, 4>(p, 2, nPairs, predSpread, st, done) : launch_smem_level<32, 0, 1, 1, 4>(p, 2, nPairs, predSpread, st, done);
  case 3: return ss ? launch_smem_level<64, 1, 1, 1, 8>(p, 3, nPairs, predSpread, st, done) : launch_smem_level<64, 0, 1, 1, 8>(p, 3, nPairs, predSpread, st, done);
  default: *done = false; return cudaSuccess;
  }
}

}   // namespace vtmme
