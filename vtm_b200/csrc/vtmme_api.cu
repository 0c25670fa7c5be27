// C ABI of libvtmme.so (include/vtmme.h): context, device pictures, and the orchestration of the kernels.
// There is no CPU implementation of anything behind these entry points: a failing CUDA call is an error.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "me_kernels.h"
#include "vtmme_internal.h"

using namespace vtmme;

struct vtmme_ctx
{
  int          device    = 0;
  cudaStream_t stream    = nullptr;
  cudaStream_t ownStream = nullptr;
  std::string  err;
  uint64_t     launches = 0;
  std::unordered_map<int, DevPic> pics;
  std::unordered_map<int, cudaEvent_t> picReady;   // pictures uploaded asynchronously on copyStream
  cudaStream_t copyStream = nullptr;

  // scratch of the frame path (grown on demand)
  DevPic*             dCur = nullptr;
  DevPic*             dRef = nullptr;
  size_t              curArrCap = 0, refArrCap = 0;
  unsigned long long* dKeys = nullptr;
  size_t              keysCap = 0;
  uint32_t*           dSurf = nullptr;
  size_t              surfCapBytes = 0;
  uint32_t*           dSurfEven = nullptr;
  size_t              surfEvenCapBytes = 0;
  int4*               dRegInfo = nullptr;
  size_t              regInfoCap = 0;
  int*                dErr = nullptr;
  int16_t*            dPred = nullptr;
  size_t              predCap = 0;
  vtmme_cu_result*    dRes = nullptr;
  size_t              resCap = 0;
  uint32_t*           dFracAcc = nullptr;
  size_t              fracAccCap = 0;

  // scratch of the per-call job path
  unsigned char* dJobBuf = nullptr;
  size_t         jobBufCap = 0;
  unsigned char* hPinned = nullptr;
  size_t         hPinnedCap = 0;
  uint32_t*      dJobSurf = nullptr;
  size_t         jobSurfCap = 0;
  unsigned long long* dJobKeys = nullptr;   // persistent, all-ones between calls (the frac kernel resets what it read)
  size_t         jobKeysCap = 0;
  uint32_t*      dJobFracAcc = nullptr;     // persistent, all-zero between calls (18 sums per job slot)
  size_t         jobFracAccCap = 0;
  unsigned int   jobSeq = 0;                // sequence number of the polled completion word (single-launch jobs)
  unsigned char* dPinnedAlias = nullptr;    // device address of hPinned (mapped, zero-copy result write-back)
  unsigned int*  dJobTicket = nullptr;      // CTA ticket of the single-launch small-job path, zero between calls
  size_t         jobTicketCap = 0;
  unsigned char* dMcTiles = nullptr;        // tile descriptors of vtmme_mc_batch / vtmme_mc_host
  size_t         mcTilesCap = 0;
  unsigned char* hMcTiles = nullptr;        // their own pinned staging block (vtmme_mc_batch returns before the upload ends)
  size_t         hMcTilesCap = 0;
  cudaEvent_t    mcUploaded = nullptr;      // recorded after the tile upload: the next call waits before re-filling

  // contiguous landing zone of the pipelined host uploads (copy stream)
  int16_t* dUpStage = nullptr;
  size_t   upStageCap = 0;

  // scratch of vtmme_mctf_me: sub-sampled planes, vector fields, picture descriptors
  unsigned char* dMctf = nullptr;
  size_t         mctfCap = 0;

  // picture descriptors of a frame call: page-locked staging, two slots, so that the call returns without
  // waiting for the copy (vtmme_search_frames_device is asynchronous)
  DevPic*     hPairs = nullptr;
  size_t      hPairsCap = 0;            // DevPic entries per slot
  cudaEvent_t pairsCopied[2] = { nullptr, nullptr };
  int         pairsSlot = 0;

  bool        profiling = false;
  cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
  bool        evValid = false;
};

int vtmme_set_error(vtmme_ctx* ctx, int code, const char* what, const char* detail)
{
  if (ctx)
  {
    ctx->err = std::string(what ? what : "") + ": " + (detail ? detail : "");
  }
  return code;
}

namespace {

template <typename T>
int ensure(vtmme_ctx* ctx, T*& ptr, size_t& capBytes, size_t needBytes)
{
  if (needBytes <= capBytes) return VTMME_OK;
  if (ptr) cudaFree(ptr);
  ptr      = nullptr;
  capBytes = 0;
  void* p  = nullptr;
  if (cudaMalloc(&p, needBytes) != cudaSuccess)
  {
    cudaGetLastError();
    return vtmme_set_error(ctx, VTMME_ERR_NOMEM, "cudaMalloc", "out of device memory");
  }
  ptr      = reinterpret_cast<T*>(p);
  capBytes = needBytes;
  return VTMME_OK;
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

int alloc_pic(vtmme_ctx* ctx, int picId, int width, int height, int margin, DevPic* out)
{
  auto it = ctx->pics.find(picId);
  const int m = round_up(margin < VTMME_MIN_MARGIN ? VTMME_MIN_MARGIN : margin, 64);
  if (it != ctx->pics.end())
  {
    if (it->second.width == width && it->second.height == height && it->second.margin == m)
    {
      *out = it->second;
      return VTMME_OK;
    }
    cudaFree(it->second.base);
    ctx->pics.erase(it);
  }
  DevPic p;
  p.width  = width;
  p.height = height;
  p.margin = m;
  p.stride = round_up(width + 2 * m, 64);
  const size_t bytes = (size_t) p.stride * (height + 2 * m) * sizeof(int16_t);
  void* base = nullptr;
  if (cudaMalloc(&base, bytes) != cudaSuccess)
  {
    cudaGetLastError();
    return vtmme_set_error(ctx, VTMME_ERR_NOMEM, "cudaMalloc(picture)", "out of device memory");
  }
  p.base   = reinterpret_cast<int16_t*>(base);
  p.origin = p.base + (size_t) m * p.stride + m;
  ctx->pics[picId] = p;
  *out = p;
  return VTMME_OK;
}

// compute-stream work that reads picture `picId` must wait for its asynchronous upload
int wait_picture(vtmme_ctx* ctx, int picId)
{
  auto it = ctx->picReady.find(picId);
  if (it == ctx->picReady.end()) return VTMME_OK;
  VTMME_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, it->second, 0));
  return VTMME_OK;
}

int upload_common(vtmme_ctx* ctx, int picId, const int16_t* origin, int stride, int width, int height, int margin,
                  int withBorder, cudaMemcpyKind kind, bool async = false)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!origin || width <= 0 || height <= 0 || stride < width || margin < 0)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_upload_picture", "bad plane description");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  DevPic p;
  int rc = alloc_pic(ctx, picId, width, height, margin, &p);
  if (rc != VTMME_OK) return rc;
  // copy what the caller owns: the picture area, plus its own border when asked to
  const int cm = withBorder ? (margin < p.margin ? margin : p.margin) : 0;
  cudaStream_t st = ctx->stream;
  if (async)
  {
    if (!ctx->copyStream)
    {
      // highest priority: the short border kernels of an upload must not queue behind the CTAs of a running search
      // (with equal priority 64 uploads took longer than the 79 ms search they overlap with)
      int lo = 0, hi = 0;
      VTMME_CUDA_CHECK(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
      VTMME_CUDA_CHECK(ctx, cudaStreamCreateWithPriority(&ctx->copyStream, cudaStreamNonBlocking, hi));
    }
    st = ctx->copyStream;
    // the copy must not overtake searches already queued on the compute stream that still read this picture
    cudaEvent_t& ev = ctx->picReady[picId];
    if (!ev) VTMME_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    VTMME_CUDA_CHECK(ctx, cudaEventRecord(ev, ctx->stream));
    VTMME_CUDA_CHECK(ctx, cudaStreamWaitEvent(st, ev, 0));
  }
  else
  {
    auto it = ctx->picReady.find(picId);   // a synchronous re-upload supersedes a pending asynchronous one
    if (it != ctx->picReady.end())
    {
      VTMME_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, it->second, 0));
      cudaEventDestroy(it->second);
      ctx->picReady.erase(it);
    }
  }
  if (async && kind == cudaMemcpyHostToDevice && cm == 0 && stride - width < 64)
  {
    // pipelined upload of a (nearly) contiguous page-locked plane: ONE transfer into a landing zone, then one kernel
    // writes the padded plane — picture area and replicated border (Picture::extendPicBorder, Picture.cpp:1050-1096) —
    // in a single pass.  (Stream order on the copy stream makes the reuse of the landing zone safe.)
    const size_t bytes = ((size_t) stride * (height - 1) + width) * 2;
    if ((rc = ensure(ctx, ctx->dUpStage, ctx->upStageCap, bytes)) != VTMME_OK) return rc;
    VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dUpStage, origin, bytes, kind, st));
    VTMME_CUDA_CHECK(ctx, launch_scatter_extend(p, ctx->dUpStage, stride, st));
  }
  else
  {
    VTMME_CUDA_CHECK(ctx, cudaMemcpy2DAsync(p.origin - (ptrdiff_t) cm * p.stride - cm, (size_t) p.stride * 2,
                                            origin - (ptrdiff_t) cm * stride - cm, (size_t) stride * 2,
                                            (size_t) (width + 2 * cm) * 2, height + 2 * cm, kind, st));
    // everything beyond is edge replication (Picture::extendPicBorder, Picture.cpp:1050-1096)
    VTMME_CUDA_CHECK(ctx, launch_extend_border(p, st));
  }
  ctx->launches += 1;
  if (async) VTMME_CUDA_CHECK(ctx, cudaEventRecord(ctx->picReady[picId], st));
  else if (kind == cudaMemcpyHostToDevice) VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return VTMME_OK;
}

}   // namespace

extern "C" {

int vtmme_create(int device, vtmme_ctx** out)
{
  if (!out) return VTMME_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return VTMME_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return VTMME_ERR_CUDA;
  vtmme_ctx* ctx = new vtmme_ctx();
  ctx->device = device;
  if (cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking) != cudaSuccess)
  {
    delete ctx;
    return VTMME_ERR_CUDA;
  }
  ctx->stream = ctx->ownStream;
  if (cudaMalloc(&ctx->dErr, sizeof(int)) != cudaSuccess)
  {
    cudaStreamDestroy(ctx->ownStream);
    delete ctx;
    return VTMME_ERR_NOMEM;
  }
  cudaMemset(ctx->dErr, 0, sizeof(int));
  *out = ctx;
  return VTMME_OK;
}

void vtmme_destroy(vtmme_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copyStream) cudaStreamSynchronize(ctx->copyStream);
  for (auto& kv : ctx->pics) cudaFree(kv.second.base);
  for (auto& kv : ctx->picReady) cudaEventDestroy(kv.second);
  if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
  cudaFree(ctx->dCur);
  cudaFree(ctx->dRef);
  cudaFree(ctx->dKeys);
  cudaFree(ctx->dSurf);
  cudaFree(ctx->dSurfEven);
  cudaFree(ctx->dRegInfo);
  cudaFree(ctx->dErr);
  cudaFree(ctx->dPred);
  cudaFree(ctx->dRes);
  cudaFree(ctx->dFracAcc);
  cudaFree(ctx->dJobBuf);
  cudaFree(ctx->dMcTiles);
  cudaFree(ctx->dJobTicket);
  if (ctx->hMcTiles) cudaFreeHost(ctx->hMcTiles);
  if (ctx->mcUploaded) cudaEventDestroy(ctx->mcUploaded);
  cudaFree(ctx->dJobSurf);
  cudaFree(ctx->dJobKeys);
  cudaFree(ctx->dJobFracAcc);
  if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
  if (ctx->hPairs) cudaFreeHost(ctx->hPairs);
  cudaFree(ctx->dUpStage);
  cudaFree(ctx->dMctf);
  for (int i = 0; i < 2; i++)
    if (ctx->pairsCopied[i]) cudaEventDestroy(ctx->pairsCopied[i]);
  for (int i = 0; i < 4; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  cudaStreamDestroy(ctx->ownStream);
  delete ctx;
}

const char* vtmme_last_error(const vtmme_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int vtmme_set_stream(vtmme_ctx* ctx, void* cudaStream)
{
  if (!ctx) return VTMME_ERR_ARG;
  ctx->stream = cudaStream ? reinterpret_cast<cudaStream_t>(cudaStream) : ctx->ownStream;
  return VTMME_OK;
}

int vtmme_synchronize(vtmme_ctx* ctx)
{
  if (!ctx) return VTMME_ERR_ARG;
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->copyStream) VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->copyStream));
  return VTMME_OK;
}

uint64_t vtmme_launch_count(const vtmme_ctx* ctx) { return ctx ? ctx->launches : 0; }

int vtmme_set_profiling(vtmme_ctx* ctx, int enable)
{
  if (!ctx) return VTMME_ERR_ARG;
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  if (enable && !ctx->ev[0])
    for (int i = 0; i < 4; i++) VTMME_CUDA_CHECK(ctx, cudaEventCreate(&ctx->ev[i]));
  ctx->profiling = enable != 0;
  ctx->evValid   = false;
  return VTMME_OK;
}

int vtmme_frame_kernel_ms(vtmme_ctx* ctx, float ms[3])
{
  if (!ctx || !ms) return VTMME_ERR_ARG;
  if (!ctx->evValid) return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_frame_kernel_ms", "no profiled frame search yet");
  VTMME_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->ev[3]));
  for (int i = 0; i < 3; i++) VTMME_CUDA_CHECK(ctx, cudaEventElapsedTime(&ms[i], ctx->ev[i], ctx->ev[i + 1]));
  return VTMME_OK;
}

int vtmme_upload_picture(vtmme_ctx* ctx, int picId, const int16_t* origin, int stride, int width, int height,
                         int margin, int withBorder)
{
  return upload_common(ctx, picId, origin, stride, width, height, margin, withBorder, cudaMemcpyHostToDevice);
}

int vtmme_upload_picture_device(vtmme_ctx* ctx, int picId, const int16_t* dOrigin, int stride, int width, int height,
                                int margin, int withBorder)
{
  return upload_common(ctx, picId, dOrigin, stride, width, height, margin, withBorder, cudaMemcpyDeviceToDevice);
}

int vtmme_upload_picture_async(vtmme_ctx* ctx, int picId, const int16_t* origin, int stride, int width, int height,
                               int margin, int withBorder)
{
  return upload_common(ctx, picId, origin, stride, width, height, margin, withBorder, cudaMemcpyHostToDevice, true);
}

int vtmme_release_picture(vtmme_ctx* ctx, int picId)
{
  if (!ctx) return VTMME_ERR_ARG;
  auto it = ctx->pics.find(picId);
  if (it == ctx->pics.end()) return vtmme_set_error(ctx, VTMME_ERR_NOPIC, "vtmme_release_picture", "unknown picture id");
  cudaStreamSynchronize(ctx->stream);
  auto ev = ctx->picReady.find(picId);
  if (ev != ctx->picReady.end())
  {
    cudaEventSynchronize(ev->second);
    cudaEventDestroy(ev->second);
    ctx->picReady.erase(ev);
  }
  cudaFree(it->second.base);
  ctx->pics.erase(it);
  return VTMME_OK;
}

int vtmme_frame_cu_count(int width, int height, int32_t levelOffset[6])
{
  if (width <= 0 || height <= 0) return VTMME_ERR_ARG;
  const FrameGeom g = make_geom(width, height);
  if (levelOffset)
    for (int i = 0; i < 6; i++) levelOffset[i] = g.off[i];
  return g.off[5];
}

int vtmme_search_frames_device(vtmme_ctx* ctx, int nPairs, const int32_t* curPics, const int32_t* refPics,
                               const vtmme_frame_params* prm, const int16_t* dPredQ, vtmme_cu_result* dResults)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (nPairs <= 0 || !curPics || !refPics || !prm || !dResults)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search_frames", "null argument");
  if (prm->searchRange < 1 || prm->searchRange > 512 || prm->bitDepth < 8 || prm->bitDepth > 10 || prm->predSpread < 0 ||
      prm->fastSearch < 0 || prm->fastSearch > 3 ||
      (prm->subShiftMode != 0 && prm->subShiftMode != 2 && !(prm->subShiftMode == 1 && prm->fastSearch == 2)) ||
      (prm->fastSearch == 2 && prm->searchRange > 128))   // tz_raster's window index (see vtmme_search)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search_frames", "unsupported parameters");
  const bool tzFrame = prm->fastSearch != 0;
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));

  // page-locked staging slot for the descriptors (the previous user of the slot is two calls back)
  if ((size_t) 2 * nPairs > ctx->hPairsCap)
  {
    VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->hPairs) cudaFreeHost(ctx->hPairs);
    ctx->hPairs    = nullptr;
    ctx->hPairsCap = 0;
    void* hp       = nullptr;
    if (cudaHostAlloc(&hp, (size_t) 2 * 2 * nPairs * sizeof(DevPic), cudaHostAllocDefault) != cudaSuccess)
    {
      cudaGetLastError();
      return vtmme_set_error(ctx, VTMME_ERR_NOMEM, "cudaHostAlloc", "out of pinned host memory");
    }
    ctx->hPairs    = reinterpret_cast<DevPic*>(hp);
    ctx->hPairsCap = (size_t) 2 * nPairs;
  }
  const int slot = ctx->pairsSlot;
  ctx->pairsSlot ^= 1;
  if (!ctx->pairsCopied[slot])
    VTMME_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->pairsCopied[slot], cudaEventDisableTiming));
  else
    VTMME_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->pairsCopied[slot]));
  DevPic* hc = ctx->hPairs + (size_t) slot * ctx->hPairsCap;
  DevPic* hr = hc + nPairs;
  for (int i = 0; i < nPairs; i++)
  {
    auto a = ctx->pics.find(curPics[i]), b = ctx->pics.find(refPics[i]);
    if (a == ctx->pics.end() || b == ctx->pics.end())
      return vtmme_set_error(ctx, VTMME_ERR_NOPIC, "vtmme_search_frames", "unknown picture id");
    hc[i] = a->second;
    hr[i] = b->second;
    int wrc;
    if ((wrc = wait_picture(ctx, curPics[i])) != VTMME_OK || (wrc = wait_picture(ctx, refPics[i])) != VTMME_OK) return wrc;
    if (hc[i].width != hc[0].width || hc[i].height != hc[0].height || hr[i].width != hc[0].width ||
        hr[i].height != hc[0].height)
      return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search_frames", "all pictures must have the same size");
  }
  const FrameGeom g    = make_geom(hc[0].width, hc[0].height);
  const int       nCU  = g.off[5];
  const int       nReg = g.nRegX * g.nRegY;
  if (nCU == 0) return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search_frames", "picture smaller than 8x8");

  // capacity of a region's displacement superset: own window + predictor spread, 8-aligned on the left
  const int spanMax = 2 * prm->searchRange + 1 + prm->predSpread;
  const int maxGx   = (spanMax + 7 + 7) / 8;
  const int maxRows = spanMax;
  // The MV clip (xClipMv) keeps every block within ctu+7 samples of the picture, so with ctu <= 128 all window,
  // staging-pad and filter-tap reads stay inside VTMME_MIN_MARGIN whatever the search range is.
  if (prm->ctuSize < 8 || prm->ctuSize > 128)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search_frames", "ctuSize must be in [8,128]");
  const int bandRows = tree_pick_band_rows(maxGx, maxRows, prm->subShiftMode == 2);
  const size_t surfCap = (size_t) maxGx * 8 * maxRows;

  int rc;
  if ((rc = ensure(ctx, ctx->dCur, ctx->curArrCap, (size_t) nPairs * sizeof(DevPic))) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dRef, ctx->refArrCap, (size_t) nPairs * sizeof(DevPic))) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dKeys, ctx->keysCap, (size_t) nPairs * nCU * 8)) != VTMME_OK) return rc;
  if (!tzFrame && (rc = ensure(ctx, ctx->dSurf, ctx->surfCapBytes, (size_t) nPairs * nReg * surfCap * 4)) != VTMME_OK) return rc;
  if (!tzFrame && prm->subShiftMode == 2 &&
      (rc = ensure(ctx, ctx->dSurfEven, ctx->surfEvenCapBytes, (size_t) nPairs * nReg * surfCap * 4)) != VTMME_OK)
    return rc;
  if ((rc = ensure(ctx, ctx->dRegInfo, ctx->regInfoCap, (size_t) nPairs * nReg * sizeof(int4))) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dFracAcc, ctx->fracAccCap, frac_frame_acc_bytes(g, nPairs))) != VTMME_OK) return rc;

  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dCur, hc, nPairs * sizeof(DevPic), cudaMemcpyHostToDevice, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dRef, hr, nPairs * sizeof(DevPic), cudaMemcpyHostToDevice, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaEventRecord(ctx->pairsCopied[slot], ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dKeys, 0xff, (size_t) nPairs * nCU * 8, ctx->stream));

  TreeParams tp;
  tp.g        = g;
  tp.cur      = ctx->dCur;
  tp.ref      = ctx->dRef;
  tp.predQ    = reinterpret_cast<const short2*>(dPredQ);
  tp.keys     = ctx->dKeys;
  tp.surf     = ctx->dSurf;
  tp.surfEven = prm->subShiftMode == 2 ? ctx->dSurfEven : nullptr;
  tp.subShiftMode = prm->subShiftMode;
  tp.regInfo  = ctx->dRegInfo;
  tp.errFlag  = ctx->dErr;
  tp.surfCap  = (int) surfCap;
  tp.maxGx    = maxGx;
  tp.maxRows  = maxRows;
  tp.bandRows = bandRows;
  tp.sr       = prm->searchRange;
  tp.ctu      = prm->ctuSize;
  tp.imvShift = prm->imvShift;
  tp.lambda   = prm->lambdaMotion;
  const bool prof = ctx->profiling;
  int        tzLaunches = 0;
  if (prof) VTMME_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  if (tzFrame)
  {
    // integer search = xTZSearch per CU (me_tz.cu); with profiling on its time is reported in the first slot
    TzFrameParams zp;
    zp.g               = g;
    zp.cur             = ctx->dCur;
    zp.ref             = ctx->dRef;
    zp.predQ           = reinterpret_cast<const short2*>(dPredQ);
    zp.keys            = ctx->dKeys;
    zp.sr              = prm->searchRange;
    zp.ctu             = prm->ctuSize;
    zp.imvShift        = prm->imvShift;
    zp.subShiftMode    = prm->subShiftMode;
    zp.extended        = prm->fastSearch == 3;
    zp.selective       = prm->fastSearch == 2;
    zp.firstSearchStop = prm->tzFirstSearchStop;
    zp.lambda          = prm->lambdaMotion;
    VTMME_CUDA_CHECK(ctx, launch_tz_frame(zp, nPairs, prm->predSpread, ctx->stream, &tzLaunches));
    if (prof) VTMME_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    if (prof) VTMME_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  }
  else
  {
    VTMME_CUDA_CHECK(ctx, launch_tree_sad(tp, nPairs, ctx->stream));
    if (prof) VTMME_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    VTMME_CUDA_CHECK(ctx, launch_tree_upper(tp, nPairs, ctx->stream));
    if (prof) VTMME_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  }

  FracFrameParams fp;
  fp.g        = g;
  fp.cur      = ctx->dCur;
  fp.ref      = ctx->dRef;
  fp.predQ    = reinterpret_cast<const short2*>(dPredQ);
  fp.keys     = ctx->dKeys;
  fp.results  = dResults;
  fp.bitDepth = prm->bitDepth;
  fp.imvShift = prm->imvShift;
  fp.useHad   = prm->useHad;
  fp.fracMode = prm->fracMode;
  fp.lambda   = prm->lambdaMotion;
  int fracLaunches = 0;
  VTMME_CUDA_CHECK(ctx, launch_frac_frame(fp, ctx->dFracAcc, nPairs, ctx->stream, &fracLaunches));
  if (prof)
  {
    VTMME_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    ctx->evValid = true;
  }
  ctx->launches += (tzFrame ? tzLaunches : 2) + fracLaunches;
  return VTMME_OK;
}

int vtmme_search_frames(vtmme_ctx* ctx, int nPairs, const int32_t* curPics, const int32_t* refPics,
                        const vtmme_frame_params* prm, const int16_t* predQ, vtmme_cu_result* results)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (nPairs <= 0 || !curPics || !results) return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search_frames", "null argument");
  auto a = ctx->pics.find(curPics[0]);
  if (a == ctx->pics.end()) return vtmme_set_error(ctx, VTMME_ERR_NOPIC, "vtmme_search_frames", "unknown picture id");
  const int nCU = make_geom(a->second.width, a->second.height).off[5];
  int       rc;
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  if ((rc = ensure(ctx, ctx->dRes, ctx->resCap, (size_t) nPairs * nCU * sizeof(vtmme_cu_result))) != VTMME_OK) return rc;
  const int16_t* dPred = nullptr;
  if (predQ)
  {
    if ((rc = ensure(ctx, ctx->dPred, ctx->predCap, (size_t) nPairs * nCU * 4)) != VTMME_OK) return rc;
    VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dPred, predQ, (size_t) nPairs * nCU * 4, cudaMemcpyHostToDevice, ctx->stream));
    dPred = ctx->dPred;
  }
  rc = vtmme_search_frames_device(ctx, nPairs, curPics, refPics, prm, dPred, ctx->dRes);
  if (rc != VTMME_OK) return rc;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(results, ctx->dRes, (size_t) nPairs * nCU * sizeof(vtmme_cu_result),
                                        cudaMemcpyDeviceToHost, ctx->stream));
  int flag = 0;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(&flag, ctx->dErr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (flag)
  {
    cudaMemsetAsync(ctx->dErr, 0, sizeof(int), ctx->stream);
    return vtmme_set_error(ctx, VTMME_ERR_RANGE, "vtmme_search_frames", "predictor spread exceeds params.predSpread");
  }
  return VTMME_OK;
}

}   // extern "C"

// ---- per-call jobs ---------------------------------------------------------------------------------------
static_assert(sizeof(vtmme_result) == sizeof(DevJobResult), "result layouts must match");

static int ensure_pinned(vtmme_ctx* ctx, size_t bytes)
{
  if (bytes <= ctx->hPinnedCap) return VTMME_OK;
  if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
  ctx->hPinned    = nullptr;
  ctx->hPinnedCap = 0;
  void* p = nullptr;
  bytes   = (bytes + 65535) & ~(size_t) 65535;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocMapped) != cudaSuccess)
  {
    cudaGetLastError();
    return vtmme_set_error(ctx, VTMME_ERR_NOMEM, "cudaHostAlloc", "out of pinned host memory");
  }
  void* dp = nullptr;
  if (cudaHostGetDevicePointer(&dp, p, 0) != cudaSuccess)
  {
    cudaGetLastError();
    cudaFreeHost(p);
    return vtmme_set_error(ctx, VTMME_ERR_CUDA, "cudaHostGetDevicePointer", "mapped pinned memory unavailable");
  }
  memset(p, 0, bytes);
  ctx->hPinned      = reinterpret_cast<unsigned char*>(p);
  ctx->dPinnedAlias = reinterpret_cast<unsigned char*>(dp);
  ctx->hPinnedCap   = bytes;
  return VTMME_OK;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t) 255; }

// DistParam::subShift of subShiftMode 1 (RdCost.cpp:291-309) for a power-of-two height
static inline int mode1_subshift(int h) { return h > 32 ? 4 : (h > 16 ? 3 : (h > 8 ? 2 : 1)); }

namespace {
// Completion of a single-launch job: the kernel's last store is a sequence number into mapped pinned memory, which the
// host polls — cudaStreamSynchronize costs several microseconds more per call on this platform.  VTMME_POLL=0 turns the
// polling off; a kernel that never reports (a fault) falls back to the stream wait, which returns its error.
const bool g_poll = !(getenv("VTMME_POLL") && getenv("VTMME_POLL")[0] == '0');
constexpr size_t kDoneOffset = 256;   // byte offset of the completion word in the pinned block (results start at 0)

inline cudaError_t wait_done(vtmme_ctx* ctx, unsigned int seq, size_t doneOffset = kDoneOffset)
{
  if (g_poll)
  {
    volatile unsigned int* f = reinterpret_cast<volatile unsigned int*>(ctx->hPinned + doneOffset);
    for (long spins = 0; spins < (1L << 26); spins++)
    {
      if (*f == seq)
      {
        std::atomic_thread_fence(std::memory_order_acquire);
        return cudaSuccess;
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
  return cudaStreamSynchronize(ctx->stream);
}

DevAmvr make_dev_amvr(const vtmme_job& j)
{
  DevAmvr d;
  memset(&d, 0, sizeof(d));
  if (j.fracMode != 2) return d;
  const vtmme_amvr& a = *j.amvr;
  d.imv     = a.imv;
  d.numCand = a.numCand;
  for (int i = 0; i < 2; i++)
  {
    d.candX[i]      = a.candX[i];
    d.candY[i]      = a.candY[i];
    d.mvpIdxBits[i] = a.mvpIdxBits[i];
  }
  d.mvpIdx  = a.mvpIdx;
  d.bits    = a.bits;
  d.posX    = j.x;
  d.posY    = j.y;
  d.picW    = a.picW;
  d.picH    = a.picH;
  d.maxCuW  = a.maxCuW;
  d.maxCuH  = a.maxCuH;
  d.fWeight = a.fWeight;
  return d;
}

struct SearchTiming
{
  bool   on = getenv("VTMME_TIMING") != nullptr;
  double prep = 0, enqueue = 0, wait = 0;
  long   calls = 0;
  ~SearchTiming()
  {
    if (on && calls)
      fprintf(stderr, "[vtmme] vtmme_search x%ld: prep %.2f us, enqueue %.2f us, wait %.2f us per call\n", calls,
              1e6 * prep / calls, 1e6 * enqueue / calls, 1e6 * wait / calls);
  }
} g_timing;
inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}   // namespace

extern "C" int vtmme_search(vtmme_ctx* ctx, const vtmme_job* jobs, int n, vtmme_result* results)
{
  if (!ctx) return VTMME_ERR_ARG;
  const double tStart = g_timing.on ? now_s() : 0;
  if (!jobs || !results || n <= 0) return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "null argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));

  // ---- validate, size the buffers
  size_t orgBytes = 0, surfElems = 0;
  int    maxGx = 1, maxRegions = 1, maxRows = 1, maxFracChunks = 1;
  long long totalRegions = 0;
  bool   anyMulti = false;
  const bool tzCall = jobs[0].tz != nullptr;
  int        maxPatternSamples = 0;
  for (int i = 0; i < n; i++)
  {
    const vtmme_job& j = jobs[i];
    if ((j.tz != nullptr) != tzCall) return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "TZ and full-search jobs cannot share a call");
    if (j.w * j.h > maxPatternSamples) maxPatternSamples = j.w * j.h;
    const bool pow2w = j.w >= 4 && j.w <= 128 && (j.w & (j.w - 1)) == 0;
    const bool pow2h = j.h >= 4 && j.h <= 128 && (j.h & (j.h - 1)) == 0;
    if (!pow2w || !pow2h) return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "w,h must be powers of two in [4,128]");
    if (!tzCall && (j.srRight < j.srLeft || j.srBottom < j.srTop || j.srRight - j.srLeft > 1024 || j.srBottom - j.srTop > 1024))
      return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "bad search range");
    if (j.bitDepth < 8 || j.bitDepth > 10 || j.subShift < 0 || j.subShift > 4 || (j.h >> j.subShift) < 1)
      return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "unsupported bitDepth / subShift");
    if (ctx->pics.find(j.refPic) == ctx->pics.end() || (!j.org && ctx->pics.find(j.curPic) == ctx->pics.end()))
      return vtmme_set_error(ctx, VTMME_ERR_NOPIC, "vtmme_search", "unknown picture id");
    const DevPic& rp = ctx->pics[j.refPic];
    int wrc;
    if ((wrc = wait_picture(ctx, j.refPic)) != VTMME_OK || (!j.org && (wrc = wait_picture(ctx, j.curPic)) != VTMME_OK)) return wrc;
    if (tzCall)
    {
      // every probe lies in the clip rectangle of clipMvInPic (at most maxCu + 8 samples outside the picture), which is
      // inside the device margin when the rectangle is the reference picture's own
      const vtmme_tz& t = *j.tz;
      if (t.picW != rp.width || t.picH != rp.height || t.maxCu < 1 || t.maxCu > 128 || t.nSeeds < 0 || t.nSeeds > 15 ||
          t.searchRange < 1 || t.searchRange > 512)
        return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "invalid vtmme_tz (picture size of refPic, maxCu <= 128, <= 15 seeds, range 1..512)");
      // ... and when the PU itself lies inside that picture
      if (j.x < 0 || j.y < 0 || j.x + j.w > rp.width || j.y + j.h > rp.height)
        return vtmme_set_error(ctx, VTMME_ERR_RANGE, "vtmme_search", "PU outside the reference picture");
      // the exhaustive scan of the selective search indexes its window with a 17-bit multiply-high division (tz_raster)
      if (t.selective && t.searchRange > 128)
        return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "the selective TZ search takes a searchRange of at most 128");
      if (t.stagedSad && mode1_subshift(j.h) != j.subShift)
        return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "stagedSad needs the subShift of subShiftMode 1 (RdCost.cpp:291-309)");
      if (j.org) orgBytes += align256((size_t) j.w * j.h * 2);
      const int nReg = ((j.w + 31) >> 5) * ((j.h + 31) >> 5);
      if (j.fracMode && nReg > maxFracChunks) maxFracChunks = nReg;
      totalRegions += nReg;
      if (j.fracMode < 0 || j.fracMode > 2) return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "fracMode must be 0, 1 or 2");
      if (j.fracMode == 2)
      {
        const vtmme_amvr* a = j.amvr;
        if (!a || (a->imv != 1 && a->imv != 2) || a->numCand < 1 || a->numCand > 2 || a->mvpIdx < 0 || a->mvpIdx >= a->numCand ||
            a->picW != rp.width || a->picH != rp.height || a->maxCuW < 1 || a->maxCuW > 128 || a->maxCuH < 1 || a->maxCuH > 128)
          return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "fracMode 2 needs a valid vtmme_amvr (imv 1|2, 1-2 candidates, picture size of refPic)");
      }
      continue;
    }
    // every read (window + pattern + 8-tap halo + staging pad) must stay inside the device margin
    if (j.x + j.srLeft - 24 < -rp.margin || j.x + j.w + j.srRight + 24 > rp.width + rp.margin ||
        j.y + j.srTop - 8 < -rp.margin || j.y + j.h + j.srBottom + 8 > rp.height + rp.margin)
      return vtmme_set_error(ctx, VTMME_ERR_RANGE, "vtmme_search", "search window leaves the padded reference picture");
    if (j.fracMode < 0 || j.fracMode > 2) return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "fracMode must be 0, 1 or 2");
    if (j.fracMode == 2)
    {
      // clipMvInPic keeps every probe within maxCu + 8 samples of the picture, i.e. inside the device margin, provided
      // the clip rectangle is the reference picture's own
      const vtmme_amvr* a = j.amvr;
      if (!a || (a->imv != 1 && a->imv != 2) || a->numCand < 1 || a->numCand > 2 || a->mvpIdx < 0 || a->mvpIdx >= a->numCand ||
          a->picW != rp.width || a->picH != rp.height || a->maxCuW < 1 || a->maxCuW > 128 || a->maxCuH < 1 || a->maxCuH > 128)
        return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_search", "fracMode 2 needs a valid vtmme_amvr (imv 1|2, 1-2 candidates, picture size of refPic)");
    }
    if (j.org) orgBytes += align256((size_t) j.w * j.h * 2);
    const int wl8 = j.srLeft & ~7;
    const int ngx = (j.srRight - wl8 + 8) >> 3, nrows = j.srBottom - j.srTop + 1;
    const int nReg = ((j.w + 31) >> 5) * ((j.h + 31) >> 5);
    maxGx      = ngx > maxGx ? ngx : maxGx;
    maxRegions = nReg > maxRegions ? nReg : maxRegions;
    if (j.fracMode && nReg > maxFracChunks) maxFracChunks = nReg;   // fractional chunks are 32x32 too
    maxRows    = nrows > maxRows ? nrows : maxRows;
    totalRegions += nReg;
    if (nReg > 1)
    {
      anyMulti = true;
      surfElems += (size_t) nReg * nrows * ngx * 8;
    }
  }
  // ---- one small job: single launch, descriptor + pattern as kernel parameters (me_job_fused_kernel)
  if (n == 1 && jobs[0].w <= 32 && jobs[0].h <= 32 && !tzCall)
  {
    const vtmme_job& j  = jobs[0];
    const DevPic&    rp = ctx->pics[j.refPic];
    FusedJobArgs     a;
    const int wl8 = j.srLeft & ~7, ngx = (j.srRight - wl8 + 8) >> 3, nrows = j.srBottom - j.srTop + 1;
    a.bandRows = 128 / ngx < 1 ? 1 : (128 / ngx > 32 ? 32 : 128 / ngx);   // <= one 8-wide tile per thread and pass
    a.job.w = j.w;
    a.job.h = j.h;
    a.job.l = j.srLeft;
    a.job.r = j.srRight;
    if (fused_job_smem_bytes(a) <= 30 * 1024)
    {
      int rc;
      if ((rc = ensure_pinned(ctx, 4096)) != VTMME_OK) return rc;
      if (!ctx->dJobTicket)
      {
        if ((rc = ensure(ctx, ctx->dJobTicket, ctx->jobTicketCap, 256)) != VTMME_OK) return rc;
        VTMME_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dJobTicket, 0, 256, ctx->stream));
      }
      if (!ctx->dJobKeys)
      {
        if ((rc = ensure(ctx, ctx->dJobKeys, ctx->jobKeysCap, 256 * 8)) != VTMME_OK) return rc;
        VTMME_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dJobKeys, 0xff, 256 * 8, ctx->stream));
        if ((rc = ensure(ctx, ctx->dJobFracAcc, ctx->jobFracAccCap, 256 * 18 * 4)) != VTMME_OK) return rc;
        VTMME_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dJobFracAcc, 0, 256 * 18 * 4, ctx->stream));
      }
      a.inlinePattern = j.org != nullptr;
      if (j.org)
      {
        for (int y = 0; y < j.h; y++) memcpy(a.pattern + (size_t) y * j.w, j.org + (ptrdiff_t) y * j.orgStride, (size_t) j.w * 2);
        a.job.org       = nullptr;
        a.job.orgStride = j.w;
      }
      else
      {
        const DevPic& cp = ctx->pics[j.curPic];
        a.job.org       = cp.origin + (ptrdiff_t) j.y * cp.stride + j.x;
        a.job.orgStride = cp.stride;
      }
      a.job.refAtPU   = rp.origin + (ptrdiff_t) j.y * rp.stride + j.x;
      a.job.refStride = rp.stride;
      a.job.t = j.srTop;
      a.job.b = j.srBottom;
      a.job.predQx = j.predQx;
      a.job.predQy = j.predQy;
      a.job.imvShift = j.imvShift;
      a.job.subShift = j.subShift;
      a.job.bitDepth = j.bitDepth;
      a.job.useHad = j.useHad;
      a.job.useAltHpel = j.useAltHpel;
      a.job.fracMode = j.fracMode;
      a.job.signedOrg = j.org != nullptr;
      a.job.lambda = j.lambdaMotion;
      a.job.amvr   = make_dev_amvr(j);
      a.key    = ctx->dJobKeys;
      a.ticket = ctx->dJobTicket;
      a.result = reinterpret_cast<DevJobResult*>(ctx->dPinnedAlias);
      if (++ctx->jobSeq == 0) ++ctx->jobSeq;   // 0 is the reset value of the polled word: never a sequence number
      a.seq    = ctx->jobSeq;
      a.done   = g_poll ? reinterpret_cast<unsigned int*>(ctx->dPinnedAlias + kDoneOffset) : nullptr;
      *reinterpret_cast<volatile unsigned int*>(ctx->hPinned + kDoneOffset) = 0;   // the block is shared with the batch path
      int grid = (nrows + a.bandRows - 1) / a.bandRows;
      if (grid > 4 * 148) grid = 4 * 148;
      const double tPrep = g_timing.on ? now_s() : 0;
      VTMME_CUDA_CHECK(ctx, launch_job_fused(a, grid, ctx->stream));
      ctx->launches += 1;
      const double tEnq = g_timing.on ? now_s() : 0;
      VTMME_CUDA_CHECK(ctx, wait_done(ctx, a.seq));
      memcpy(results, ctx->hPinned, sizeof(vtmme_result));
      if (g_timing.on)
      {
        const double tEnd = now_s();
        g_timing.prep += tPrep - tStart;
        g_timing.enqueue += tEnq - tPrep;
        g_timing.wait += tEnd - tEnq;
        g_timing.calls++;
      }
      return VTMME_OK;
    }
  }

  // window rows per CTA pass: 32, or fewer when the call is small, so that even one 8x8 search spreads over >= 16 SMs
  int bandRows = 32;
  while (bandRows > 8 && totalRegions * ((maxRows + bandRows - 1) / bandRows) < 148) bandRows >>= 1;
  const int maxBands = (maxRows + bandRows - 1) / bandRows;
  int nSplit = (int) ((4 * 148 + totalRegions - 1) / totalRegions);
  nSplit     = nSplit < 1 ? 1 : (nSplit > maxBands ? maxBands : nSplit);

  // ---- one pinned staging block, one device block: [jobs][surfOff][patterns] + [keys][results]
  const size_t offJobs = 0, offSurfOff = align256(offJobs + (size_t) n * sizeof(DevJob));
  const size_t offTz = align256(offSurfOff + (size_t) n * sizeof(long long));
  const size_t offOrg = align256(offTz + (tzCall ? (size_t) n * sizeof(DevTz) : 0));
  const size_t upBytes = offOrg + orgBytes;
  const size_t offKeys = align256(upBytes), offRes = align256(offKeys + (size_t) n * 8);
  const size_t offDone  = align256(offRes + (size_t) n * sizeof(DevJobResult));   // completion word of a single-launch TZ job
  const size_t devBytes = offDone + 256;
  int rc;
  if ((rc = ensure_pinned(ctx, devBytes)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, devBytes)) != VTMME_OK) return rc;
  if (anyMulti && (rc = ensure(ctx, ctx->dJobSurf, ctx->jobSurfCap, surfElems * 4)) != VTMME_OK) return rc;

  // tiny uploads (descriptors, patterns up to 16x16): the kernels read them straight from the
  // mapped pinned block, which saves the enqueue + DMA latency of a copy; anything larger is copied to HBM first (reads over PCIe are slow)
  const bool           zeroCopy = upBytes <= 2048;
  const unsigned char* dIn      = zeroCopy ? ctx->dPinnedAlias : ctx->dJobBuf;
  DevJob*    hj   = reinterpret_cast<DevJob*>(ctx->hPinned + offJobs);
  long long* hoff = reinterpret_cast<long long*>(ctx->hPinned + offSurfOff);
  size_t     orgCur = offOrg, surfCur = 0;
  for (int i = 0; i < n; i++)
  {
    const vtmme_job& j  = jobs[i];
    const DevPic&    rp = ctx->pics[j.refPic];
    DevJob           d;
    if (j.org)
    {
      int16_t* dst = reinterpret_cast<int16_t*>(ctx->hPinned + orgCur);
      for (int y = 0; y < j.h; y++) memcpy(dst + (size_t) y * j.w, j.org + (size_t) y * j.orgStride, (size_t) j.w * 2);
      d.org       = reinterpret_cast<const int16_t*>(dIn + orgCur);
      d.orgStride = j.w;
      orgCur += align256((size_t) j.w * j.h * 2);
    }
    else
    {
      const DevPic& cp = ctx->pics[j.curPic];
      d.org       = cp.origin + (ptrdiff_t) j.y * cp.stride + j.x;
      d.orgStride = cp.stride;
    }
    d.refAtPU   = rp.origin + (ptrdiff_t) j.y * rp.stride + j.x;
    d.refStride = rp.stride;
    d.w = j.w;
    d.h = j.h;
    d.l = j.srLeft;
    d.r = j.srRight;
    d.t = j.srTop;
    d.b = j.srBottom;
    d.predQx = j.predQx;
    d.predQy = j.predQy;
    d.imvShift = j.imvShift;
    d.subShift = j.subShift;
    d.bitDepth = j.bitDepth;
    d.useHad = j.useHad;
    d.useAltHpel = j.useAltHpel;
    d.fracMode = j.fracMode;
    d.signedOrg = j.org != nullptr;
    d.lambda = j.lambdaMotion;
    d.amvr   = make_dev_amvr(j);
    if (tzCall)
    {
      const vtmme_tz& t = *j.tz;
      DevTz&          z = reinterpret_cast<DevTz*>(ctx->hPinned + offTz)[i];
      z.startX      = t.startX;
      z.startY      = t.startY;
      z.hasInt2Nx2N = t.hasInt2Nx2N;
      z.int2Nx2NX   = t.int2Nx2NX;
      z.int2Nx2NY   = t.int2Nx2NY;
      z.nSeeds      = t.nSeeds;
      for (int k = 0; k < 16; k++)
      {
        z.seedX[k] = t.seedX[k];
        z.seedY[k] = t.seedY[k];
      }
      z.searchRange     = t.searchRange;
      z.extended        = t.extended;
      z.fast            = t.fast;
      z.firstSearchStop = t.firstSearchStop;
      z.posX            = j.x;
      z.posY            = j.y;
      z.picW            = t.picW;
      z.picH            = t.picH;
      z.maxCuW = z.maxCuH = t.maxCu;
      z.selective = t.selective != 0;
      z.staged    = t.stagedSad != 0;
      d.l = d.r = d.t = d.b = 0;
    }
    hj[i]   = d;
    hoff[i] = (long long) surfCur;
    const int nReg = ((j.w + 31) >> 5) * ((j.h + 31) >> 5);
    if (nReg > 1 && !tzCall)
    {
      const int wl8 = j.srLeft & ~7;
      surfCur += (size_t) nReg * (j.srBottom - j.srTop + 1) * (((j.srRight - wl8 + 8) >> 3) * 8);
    }
  }
  const double tPrep = g_timing.on ? now_s() : 0;
  if (!zeroCopy) VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobBuf, ctx->hPinned, upBytes, cudaMemcpyHostToDevice, ctx->stream));
  if ((size_t) n * 8 > ctx->jobKeysCap)
  {
    const size_t want = (size_t) (n < 256 ? 256 : n) * 8;
    if ((rc = ensure(ctx, ctx->dJobKeys, ctx->jobKeysCap, want)) != VTMME_OK) return rc;
    VTMME_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dJobKeys, 0xff, want, ctx->stream));
    if ((rc = ensure(ctx, ctx->dJobFracAcc, ctx->jobFracAccCap, want / 8 * 18 * 4)) != VTMME_OK) return rc;
    VTMME_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->dJobFracAcc, 0, want / 8 * 18 * 4, ctx->stream));
  }
  int                launches = 0;
  bool               fusedTz  = false;
  if (++ctx->jobSeq == 0) ++ctx->jobSeq;   // 0 is the reset value of the polled word
  const unsigned int seq      = ctx->jobSeq;
  *reinterpret_cast<volatile unsigned int*>(ctx->hPinned + offDone) = 0;
  VTMME_CUDA_CHECK(ctx, launch_job_search_impl(reinterpret_cast<const DevJob*>(dIn + offJobs),
                                               ctx->dJobKeys,
                                               reinterpret_cast<DevJobResult*>(ctx->dPinnedAlias + offRes), n, maxRegions,
                                               nSplit, bandRows, maxGx, anyMulti, ctx->dJobSurf,
                                               reinterpret_cast<const long long*>(dIn + offSurfOff),
                                               ctx->dJobFracAcc, maxFracChunks, ctx->stream, &launches,
                                               tzCall ? reinterpret_cast<const DevTz*>(dIn + offTz) : nullptr,
                                               maxPatternSamples,
                                               g_poll ? reinterpret_cast<unsigned int*>(ctx->dPinnedAlias + offDone) : nullptr, seq,
                                               &fusedTz));
  ctx->launches += launches;
  // the frac kernel wrote the results straight into the mapped pinned block: no device-to-host copy
  const double tEnq = g_timing.on ? now_s() : 0;
  if (fusedTz)
    VTMME_CUDA_CHECK(ctx, wait_done(ctx, seq, offDone));
  else
    VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(results, ctx->hPinned + offRes, (size_t) n * sizeof(vtmme_result));
  if (g_timing.on)
  {
    const double tEnd = now_s();
    g_timing.prep += tPrep - tStart;
    g_timing.enqueue += tEnq - tPrep;
    g_timing.wait += tEnd - tEnq;
    g_timing.calls++;
  }
  return VTMME_OK;
}

// ---- GOP-based temporal filter: motion estimation -------------------------------------------------------------------
extern "C" int vtmme_mctf_me(vtmme_ctx* ctx, int nPairs, const int32_t* orgPics, const int32_t* refPics, int bitDepth,
                             int32_t* mv)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (nPairs <= 0 || !orgPics || !refPics || !mv || bitDepth < 8 || bitDepth > 12)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_mctf_me", "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  std::vector<DevPic> pics((size_t) 6 * nPairs);   // [level 0,1,2][org, ref][pair]
  int W = 0, H = 0;
  for (int i = 0; i < nPairs; i++)
  {
    auto a = ctx->pics.find(orgPics[i]), b = ctx->pics.find(refPics[i]);
    if (a == ctx->pics.end() || b == ctx->pics.end())
      return vtmme_set_error(ctx, VTMME_ERR_NOPIC, "vtmme_mctf_me", "unknown picture id");
    if (i == 0)
    {
      W = a->second.width;
      H = a->second.height;
    }
    if (a->second.width != W || a->second.height != H || b->second.width != W || b->second.height != H)
      return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_mctf_me", "all pictures must have the same size");
    int wrc;
    if ((wrc = wait_picture(ctx, orgPics[i])) != VTMME_OK || (wrc = wait_picture(ctx, refPics[i])) != VTMME_OK) return wrc;
    pics[(size_t) 0 * nPairs + i] = a->second;
    pics[(size_t) 1 * nPairs + i] = b->second;
  }
  if (W < 32 || H < 32) return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_mctf_me", "picture smaller than 32x32");
  // scratch layout: descriptors | sub-sampled planes (margin 128, the filter's own padding) | vector fields
  const int    pad = 128;
  const int    w2 = W / 2, h2 = H / 2, w4 = w2 / 2, h4 = h2 / 2;
  const int    s2 = round_up(w2 + 2 * pad, 64), s4 = round_up(w4 + 2 * pad, 64);
  const size_t plane2 = align256((size_t) s2 * (h2 + 2 * pad) * 2), plane4 = align256((size_t) s4 * (h4 + 2 * pad) * 2);
  const int    lw = W / 16, lh = H / 16, fw = W / 4, fh = H / 4;
  const size_t offDesc = 0, offPlanes = align256(pics.size() * sizeof(DevPic));
  const size_t offMv = offPlanes + (size_t) nPairs * 2 * (plane2 + plane4);
  // (the four vector fields are contiguous, unpadded: one kernel initialises them as a single int3 array)
  const size_t lowBytes = (size_t) nPairs * lw * lh * sizeof(int3), finBytes = (size_t) nPairs * fw * fh * sizeof(int3);
  const size_t total = offMv + 3 * lowBytes + finBytes;
  int rc;
  if ((rc = ensure(ctx, ctx->dMctf, ctx->mctfCap, total)) != VTMME_OK) return rc;
  auto sub = [&](size_t off, int w, int h, int stride) {
    DevPic p;
    p.base   = reinterpret_cast<int16_t*>(ctx->dMctf + off);
    p.stride = stride;
    p.width  = w;
    p.height = h;
    p.margin = pad;
    p.origin = p.base + (size_t) pad * stride + pad;
    return p;
  };
  for (int i = 0; i < nPairs; i++)
    for (int k = 0; k < 2; k++)   // 0 original, 1 reference
    {
      const size_t o = offPlanes + ((size_t) i * 2 + k) * (plane2 + plane4);
      pics[(size_t) (2 + k) * nPairs + i] = sub(o, w2, h2, s2);
      pics[(size_t) (4 + k) * nPairs + i] = sub(o + plane2, w4, h4, s4);
      VTMME_CUDA_CHECK(ctx, launch_mctf_subsample(pics[(size_t) k * nPairs + i], pics[(size_t) (2 + k) * nPairs + i], ctx->stream));
      VTMME_CUDA_CHECK(ctx, launch_mctf_subsample(pics[(size_t) (2 + k) * nPairs + i], pics[(size_t) (4 + k) * nPairs + i], ctx->stream));
      ctx->launches += 2;
    }
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dMctf + offDesc, pics.data(), pics.size() * sizeof(DevPic), cudaMemcpyHostToDevice, ctx->stream));
  int3* mv0  = reinterpret_cast<int3*>(ctx->dMctf + offMv);
  int3* mv1  = reinterpret_cast<int3*>(ctx->dMctf + offMv + lowBytes);
  int3* mv2  = reinterpret_cast<int3*>(ctx->dMctf + offMv + 2 * lowBytes);
  int3* mvF  = reinterpret_cast<int3*>(ctx->dMctf + offMv + 3 * lowBytes);
  VTMME_CUDA_CHECK(ctx, launch_mctf_init_mv(mv0, (int) ((3 * lowBytes + finBytes) / sizeof(int3)), ctx->stream));
  const DevPic* d = reinterpret_cast<const DevPic*>(ctx->dMctf + offDesc);
  MctfLevelParams p;
  p.maxv = (1 << bitDepth) - 1;
  // motionEstimationLuma(mv_0, origSubsampled4, bufferSub4, 16)                      (:461)
  p.org = d + (size_t) 4 * nPairs;  p.ref = d + (size_t) 5 * nPairs;  p.width = w4;  p.height = h4;
  p.previous = nullptr;  p.prevW = p.prevH = 0;  p.factor = 1;  p.mvs = mv0;  p.mvW = lw;  p.mvH = lh;
  VTMME_CUDA_CHECK(ctx, launch_mctf_level(p, 16, false, nPairs, ctx->stream));
  // motionEstimationLuma(mv_1, origSubsampled2, bufferSub2, 16, &mv_0, 2)             (:462)
  p.org = d + (size_t) 2 * nPairs;  p.ref = d + (size_t) 3 * nPairs;  p.width = w2;  p.height = h2;
  p.previous = mv0;  p.prevW = lw;  p.prevH = lh;  p.factor = 2;  p.mvs = mv1;
  VTMME_CUDA_CHECK(ctx, launch_mctf_level(p, 16, false, nPairs, ctx->stream));
  // motionEstimationLuma(mv_2, orgPic, buffer, 16, &mv_1, 2)                          (:463)
  p.org = d;  p.ref = d + (size_t) nPairs;  p.width = W;  p.height = H;
  p.previous = mv1;  p.mvs = mv2;
  VTMME_CUDA_CHECK(ctx, launch_mctf_level(p, 16, false, nPairs, ctx->stream));
  // motionEstimationLuma(mv, orgPic, buffer, 8, &mv_2, 1, true)                       (:465)
  p.previous = mv2;  p.factor = 1;  p.mvs = mvF;  p.mvW = fw;  p.mvH = fh;
  VTMME_CUDA_CHECK(ctx, launch_mctf_level(p, 8, true, nPairs, ctx->stream));
  ctx->launches += 5;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(mv, mvF, finBytes, cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return VTMME_OK;
}

extern "C" int vtmme_mctf_apply_motion(vtmme_ctx* ctx, int srcPic, int csx, int csy, const int32_t* mv, int mvStride,
                                       int mvRows, int bitDepth, int16_t* dst)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!mv || !dst || csx < 0 || csx > 1 || csy < 0 || csy > 1 || mvStride <= 0 || mvRows <= 0 || bitDepth < 8 || bitDepth > 12)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_mctf_apply_motion", "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  auto it = ctx->pics.find(srcPic);
  if (it == ctx->pics.end()) return vtmme_set_error(ctx, VTMME_ERR_NOPIC, "vtmme_mctf_apply_motion", "unknown picture id");
  int rc;
  if ((rc = wait_picture(ctx, srcPic)) != VTMME_OK) return rc;
  const DevPic& sp = it->second;
  const int bsx = 8 >> csx, bsy = 8 >> csy;
  if (sp.width / bsx > mvStride || sp.height / bsy > mvRows)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_mctf_apply_motion", "vector field smaller than the block grid");
  // vectors are bounded by the pyramid's ranges; anything that would leave the padded plane is refused
  const int lim = (sp.margin - 4) << 4;
  for (int y = 0; y < sp.height / bsy; y++)
    for (int x = 0; x < sp.width / bsx; x++)
    {
      const int32_t* m = mv + 3 * ((size_t) y * mvStride + x);
      if ((m[0] >> csx) > lim || (m[0] >> csx) < -lim || (m[1] >> csy) > lim || (m[1] >> csy) < -lim)
        return vtmme_set_error(ctx, VTMME_ERR_RANGE, "vtmme_mctf_apply_motion", "motion vector leaves the padded plane");
    }
  const size_t mvBytes = (size_t) mvStride * mvRows * sizeof(int3), outBytes = (size_t) sp.width * sp.height * 2;
  if ((rc = ensure(ctx, ctx->dMctf, ctx->mctfCap, align256(mvBytes) + outBytes)) != VTMME_OK) return rc;
  int3*    dMv  = reinterpret_cast<int3*>(ctx->dMctf);
  int16_t* dOut = reinterpret_cast<int16_t*>(ctx->dMctf + align256(mvBytes));
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(dMv, mv, mvBytes, cudaMemcpyHostToDevice, ctx->stream));
  VTMME_CUDA_CHECK(ctx, launch_mctf_apply_motion(sp, csx, csy, dMv, mvStride, (1 << bitDepth) - 1, dOut, ctx->stream));
  ctx->launches += 1;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, dOut, outBytes, cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return VTMME_OK;
}

extern "C" int vtmme_mctf_bilateral(vtmme_ctx* ctx, int orgPic, int numRefs, const int32_t* corrPics, const double* weights,
                                    int bitDepth, int16_t* dst)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (numRefs < 1 || numRefs > 8 || !corrPics || !weights || !dst || bitDepth < 8 || bitDepth > 12)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_mctf_bilateral", "bad argument (1..8 neighbouring pictures, bit depth 8..12)");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  auto it = ctx->pics.find(orgPic);
  if (it == ctx->pics.end()) return vtmme_set_error(ctx, VTMME_ERR_NOPIC, "vtmme_mctf_bilateral", "unknown picture id");
  int rc;
  if ((rc = wait_picture(ctx, orgPic)) != VTMME_OK) return rc;
  const DevPic op = it->second;
  DevPic       corr[8];
  for (int i = 0; i < numRefs; i++)
  {
    auto c = ctx->pics.find(corrPics[i]);
    if (c == ctx->pics.end()) return vtmme_set_error(ctx, VTMME_ERR_NOPIC, "vtmme_mctf_bilateral", "unknown picture id");
    if (c->second.width != op.width || c->second.height != op.height)
      return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_mctf_bilateral", "pictures differ in size");
    if ((rc = wait_picture(ctx, corrPics[i])) != VTMME_OK) return rc;
    corr[i] = c->second;
  }
  const size_t tabBytes = (size_t) numRefs * ((size_t) 1 << bitDepth) * sizeof(double), outBytes = (size_t) op.width * op.height * 2;
  if ((rc = ensure(ctx, ctx->dMctf, ctx->mctfCap, align256(tabBytes) + outBytes)) != VTMME_OK) return rc;
  double*  dTab = reinterpret_cast<double*>(ctx->dMctf);
  int16_t* dOut = reinterpret_cast<int16_t*>(ctx->dMctf + align256(tabBytes));
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(dTab, weights, tabBytes, cudaMemcpyHostToDevice, ctx->stream));
  VTMME_CUDA_CHECK(ctx, launch_mctf_bilateral(op, corr, numRefs, dTab, bitDepth, dOut, ctx->stream));
  ctx->launches += 1;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, dOut, outBytes, cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return VTMME_OK;
}

// ---- table-level entry points -------------------------------------------------------------------------------
extern "C" int vtmme_dist_batch(vtmme_ctx* ctx, int kind, const int16_t* dOrg, int orgStride, int64_t orgBlockStride,
                                const int16_t* dCur, int curStride, int64_t curBlockStride, int w, int h, int subShift,
                                int n, uint64_t* dOut)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!dOrg || !dCur || !dOut || n <= 0 || w < 2 || h < 2 || w > 128 || h > 128 || (w & 1) || (h & 1) || kind < 0 ||
      kind > 1 || subShift < 0 || (h >> subShift) < 1)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_dist_batch", "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  VTMME_CUDA_CHECK(ctx, launch_dist_batch(kind, dOrg, orgStride, orgBlockStride, dCur, curStride, curBlockStride, w, h,
                                          kind == 0 ? subShift : 0, n, reinterpret_cast<unsigned long long*>(dOut),
                                          ctx->stream));
  ctx->launches += 1;
  return VTMME_OK;
}

extern "C" int vtmme_dist_host(vtmme_ctx* ctx, int kind, const int16_t* org, int orgStride, const int16_t* cur,
                               int curStride, int w, int h, int subShift, uint64_t* out)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!org || !cur || !out || w < 2 || h < 2 || w > 128 || h > 128)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_dist_host", "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const size_t blk = align256((size_t) w * h * 2);
  int          rc;
  if ((rc = ensure_pinned(ctx, 2 * blk + 256)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, 2 * blk + 256)) != VTMME_OK) return rc;
  int16_t* ho = reinterpret_cast<int16_t*>(ctx->hPinned);
  int16_t* hc = reinterpret_cast<int16_t*>(ctx->hPinned + blk);
  for (int y = 0; y < h; y++)
  {
    memcpy(ho + (size_t) y * w, org + (size_t) y * orgStride, (size_t) w * 2);
    memcpy(hc + (size_t) y * w, cur + (size_t) y * curStride, (size_t) w * 2);
  }
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobBuf, ctx->hPinned, 2 * blk, cudaMemcpyHostToDevice, ctx->stream));
  rc = vtmme_dist_batch(ctx, kind, reinterpret_cast<int16_t*>(ctx->dJobBuf), w, 0,
                        reinterpret_cast<int16_t*>(ctx->dJobBuf + blk), w, 0, w, h, subShift, 1,
                        reinterpret_cast<uint64_t*>(ctx->dJobBuf + 2 * blk));
  if (rc != VTMME_OK) return rc;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->hPinned + 2 * blk, ctx->dJobBuf + 2 * blk, 8, cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  *out = *reinterpret_cast<uint64_t*>(ctx->hPinned + 2 * blk);
  return VTMME_OK;
}

static int interp_args_ok(int comp, int w, int h, int frac, int bitDepth)
{
  if (comp < 0 || comp > 1 || w < 1 || h < 1 || w > 256 || h > 256 || bitDepth < 8 || bitDepth > 10) return 0;
  if (frac < 0 || frac >= (comp == 0 ? 16 : 32)) return 0;
  return 1;
}

extern "C" int vtmme_interp_batch(vtmme_ctx* ctx, int comp, int vertical, const int16_t* dSrc, int srcStride,
                                  int64_t srcBlockStride, int16_t* dDst, int dstStride, int64_t dstBlockStride, int w,
                                  int h, int frac, int isFirst, int isLast, int bitDepth, int useAltHpel, int n)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!dSrc || !dDst || n <= 0 || !interp_args_ok(comp, w, h, frac, bitDepth))
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_interp_batch", "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  VTMME_CUDA_CHECK(ctx, launch_interp_batch(comp, vertical, dSrc, srcStride, srcBlockStride, dDst, dstStride,
                                            dstBlockStride, w, h, frac, isFirst, isLast, bitDepth, useAltHpel, n,
                                            ctx->stream));
  ctx->launches += 1;
  return VTMME_OK;
}

extern "C" int vtmme_interp_host(vtmme_ctx* ctx, int comp, int vertical, const int16_t* src, int srcStride, int16_t* dst,
                                 int dstStride, int w, int h, int frac, int isFirst, int isLast, int bitDepth,
                                 int useAltHpel)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!src || !dst || !interp_args_ok(comp, w, h, frac, bitDepth))
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_interp_host", "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  // source halo: (N/2-1) samples before and N/2 after along the filtered direction
  const int taps = comp == 0 ? 8 : 4, before = taps / 2 - 1, after = taps / 2;
  const int sw = w + (vertical ? 0 : before + after), sh = h + (vertical ? before + after : 0);
  const size_t srcBytes = align256((size_t) sw * sh * 2), dstBytes = align256((size_t) w * h * 2);
  int rc;
  if ((rc = ensure_pinned(ctx, srcBytes + dstBytes)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, srcBytes + dstBytes)) != VTMME_OK) return rc;
  int16_t*       hs = reinterpret_cast<int16_t*>(ctx->hPinned);
  const int16_t* s0 = src - (vertical ? (ptrdiff_t) before * srcStride : before);
  for (int y = 0; y < sh; y++) memcpy(hs + (size_t) y * sw, s0 + (ptrdiff_t) y * srcStride, (size_t) sw * 2);
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobBuf, ctx->hPinned, srcBytes, cudaMemcpyHostToDevice, ctx->stream));
  const int16_t* dsrc = reinterpret_cast<int16_t*>(ctx->dJobBuf) + (vertical ? (size_t) before * sw : before);
  rc = vtmme_interp_batch(ctx, comp, vertical, dsrc, sw, 0, reinterpret_cast<int16_t*>(ctx->dJobBuf + srcBytes), w, 0, w, h,
                          frac, isFirst, isLast, bitDepth, useAltHpel, 1);
  if (rc != VTMME_OK) return rc;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->hPinned + srcBytes, ctx->dJobBuf + srcBytes, (size_t) w * h * 2,
                                        cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const int16_t* hd = reinterpret_cast<int16_t*>(ctx->hPinned + srcBytes);
  for (int y = 0; y < h; y++) memcpy(dst + (size_t) y * dstStride, hd + (size_t) y * w, (size_t) w * 2);
  return VTMME_OK;
}

extern "C" int vtmme_filter_host(vtmme_ctx* ctx, int nTaps, int vertical, int isFirst, int isLast, int copy, const int16_t* src,
                                 int srcStride, int16_t* dst, int dstStride, int w, int h, const int16_t* coeff, int bitDepth)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!src || !dst || (!copy && !coeff) || (nTaps != 8 && nTaps != 4 && nTaps != 2) || w < 1 || h < 1 || w > 256 || h > 256 ||
      bitDepth < 8 || bitDepth > 10)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_filter_host", "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const int before = copy ? 0 : nTaps / 2 - 1, after = copy ? 0 : nTaps / 2;
  const int sw = w + (vertical ? 0 : before + after), sh = h + (vertical ? before + after : 0);
  const size_t srcBytes = align256((size_t) sw * sh * 2), dstBytes = align256((size_t) w * h * 2);
  int rc;
  if ((rc = ensure_pinned(ctx, srcBytes + dstBytes)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, srcBytes + dstBytes)) != VTMME_OK) return rc;
  int16_t*       hs = reinterpret_cast<int16_t*>(ctx->hPinned);
  const int16_t* s0 = src - (vertical ? (ptrdiff_t) before * srcStride : before);
  for (int y = 0; y < sh; y++) memcpy(hs + (size_t) y * sw, s0 + (ptrdiff_t) y * srcStride, (size_t) sw * 2);
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobBuf, ctx->hPinned, srcBytes, cudaMemcpyHostToDevice, ctx->stream));
  const int16_t* dsrc = reinterpret_cast<int16_t*>(ctx->dJobBuf) + (vertical ? (size_t) before * sw : before);
  VTMME_CUDA_CHECK(ctx, launch_filter_batch(nTaps, vertical, isFirst, isLast, copy, dsrc, sw, 0,
                                            reinterpret_cast<int16_t*>(ctx->dJobBuf + srcBytes), w, 0, w, h, coeff, bitDepth,
                                            1, ctx->stream));
  ctx->launches += 1;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->hPinned + srcBytes, ctx->dJobBuf + srcBytes, (size_t) w * h * 2,
                                        cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const int16_t* hd = reinterpret_cast<int16_t*>(ctx->hPinned + srcBytes);
  for (int y = 0; y < h; y++) memcpy(dst + (size_t) y * dstStride, hd + (size_t) y * w, (size_t) w * 2);
  return VTMME_OK;
}

// ---- motion compensation -------------------------------------------------------------------------------------------
// Validates the blocks, cuts them into <=16x16 tiles in the pinned staging block, uploads the tiles to *dTiles (device
// scratch) and launches the kernel writing to dDst (block i packed at the running sum of w*h).
// page-locked staging of nTiles tile descriptors (waits until the previous upload out of it has completed)
static int mc_tiles_begin(vtmme_ctx* ctx, int nTiles)
{
  const size_t tileBytes = align256((size_t) nTiles * sizeof(McTile));
  int rc;
  if (!ctx->mcUploaded) VTMME_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->mcUploaded, cudaEventDisableTiming));
  else VTMME_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->mcUploaded));
  if (tileBytes > ctx->hMcTilesCap)
  {
    if (ctx->hMcTiles) cudaFreeHost(ctx->hMcTiles);
    ctx->hMcTiles    = nullptr;
    ctx->hMcTilesCap = 0;
    void* p = nullptr;
    if (cudaHostAlloc(&p, tileBytes, cudaHostAllocDefault) != cudaSuccess)
    {
      cudaGetLastError();
      return vtmme_set_error(ctx, VTMME_ERR_NOMEM, "cudaHostAlloc", "out of pinned host memory");
    }
    ctx->hMcTiles    = reinterpret_cast<unsigned char*>(p);
    ctx->hMcTilesCap = tileBytes;
  }
  if (tileBytes > ctx->mcTilesCap)
  {
    // the previous launch may still read the old descriptor buffer
    VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = ensure(ctx, ctx->dMcTiles, ctx->mcTilesCap, tileBytes)) != VTMME_OK) return rc;
  }
  return VTMME_OK;
}

// uploads the staged descriptors and launches the kernel
static int mc_tiles_launch(vtmme_ctx* ctx, int comp, int nTiles, int bi, int bitDepth, int useAltHpel)
{
  McTile* dTiles = reinterpret_cast<McTile*>(ctx->dMcTiles);
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(dTiles, ctx->hMcTiles, (size_t) nTiles * sizeof(McTile), cudaMemcpyHostToDevice, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaEventRecord(ctx->mcUploaded, ctx->stream));
  VTMME_CUDA_CHECK(ctx, launch_mc_batch(comp, dTiles, nTiles, bi, bitDepth, useAltHpel, ctx->stream));
  ctx->launches += 1;
  return VTMME_OK;
}

static int mc_launch(vtmme_ctx* ctx, const char* who, int comp, int bi, int bitDepth, int useAltHpel, int n,
                     const vtmme_mc_block* blocks, int nTiles, int16_t* dDst)
{
  int rc;
  if ((rc = mc_tiles_begin(ctx, nTiles)) != VTMME_OK) return rc;
  const int shift = 4 + (comp ? 1 : 0), taps = comp ? 4 : 8, before = taps / 2 - 1, after = taps / 2;
  McTile* ht = reinterpret_cast<McTile*>(ctx->hMcTiles);
  int     k = 0;
  size_t  off = 0;
  for (int i = 0; i < n; i++)
  {
    const vtmme_mc_block& b = blocks[i];
    const DevPic& rp = ctx->pics[b.refPic];
    int wrc;
    if ((wrc = wait_picture(ctx, b.refPic)) != VTMME_OK) return wrc;
    const int ix = b.mvX >> shift, iy = b.mvY >> shift, xFrac = b.mvX & ((1 << shift) - 1), yFrac = b.mvY & ((1 << shift) - 1);
    if (b.x + ix - before < -rp.margin || b.x + b.w + ix + after > rp.width + rp.margin ||
        b.y + iy - before < -rp.margin || b.y + b.h + iy + after > rp.height + rp.margin)
      return vtmme_set_error(ctx, VTMME_ERR_RANGE, who, "block + MV leaves the padded reference plane (clip the MV first)");
    // 4x4 coefficient table: keyed on the (w,h) the reference hands each pass; the horizontal pass of the two-stage
    // case runs over h + taps - 1 rows (InterpolationFilter.cpp:786, 869; InterPrediction.cpp:763)
    const int  hHor  = yFrac ? b.h + taps - 1 : b.h;
    const bool q4Hor = comp == 0 && b.w == 4 && (hHor == 4 || hHor == 11), q4Ver = comp == 0 && b.w == 4 && b.h == 4;
    const int16_t* src = rp.origin + (ptrdiff_t) (b.y + iy) * rp.stride + b.x + ix;
    for (int ty = 0; ty < b.h; ty += 16)
      for (int tx = 0; tx < b.w; tx += 16)
      {
        McTile& t = ht[k++];
        t.src = src + (ptrdiff_t) ty * rp.stride + tx;
        t.dst = dDst + off + (size_t) ty * b.w + tx;
        t.srcStride = rp.stride;
        t.dstStride = b.w;
        t.tw = (uint8_t) (b.w - tx < 16 ? b.w - tx : 16);
        t.th = (uint8_t) (b.h - ty < 16 ? b.h - ty : 16);
        t.xFrac = (uint8_t) xFrac;
        t.yFrac = (uint8_t) yFrac;
        t.q4Hor = q4Hor;
        t.q4Ver = q4Ver;
        t.winMaxX = t.winMaxY = 0;
        t.winX = t.winY = 0;
      }
    off += (size_t) b.w * b.h;
  }
  if (k != nTiles) return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "internal: tile count");
  return mc_tiles_launch(ctx, comp, nTiles, bi, bitDepth, useAltHpel);
}

// shared argument check; returns the tile count and the packed output size
static int mc_check(vtmme_ctx* ctx, const char* who, int comp, int bitDepth, int n, const vtmme_mc_block* blocks, const void* dst,
                    int* nTiles, size_t* outElems)
{
  if (!blocks || !dst || n <= 0 || (comp != 0 && comp != 1) || bitDepth < 8 || bitDepth > 10)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "bad argument");
  long long tiles = 0;
  size_t    elems = 0;
  for (int i = 0; i < n; i++)
  {
    const vtmme_mc_block& b = blocks[i];
    if (b.w < 1 || b.h < 1 || b.w > 128 || b.h > 128) return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "w,h must be in [1,128]");
    if (ctx->pics.find(b.refPic) == ctx->pics.end()) return vtmme_set_error(ctx, VTMME_ERR_NOPIC, who, "unknown picture id");
    tiles += (long long) ((b.w + 15) / 16) * ((b.h + 15) / 16);
    elems += (size_t) b.w * b.h;
  }
  if (tiles > (1 << 24)) return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "too many blocks in one call");
  *nTiles   = (int) tiles;
  *outElems = elems;
  return VTMME_OK;
}

extern "C" int vtmme_mc_batch(vtmme_ctx* ctx, int comp, int bi, int bitDepth, int useAltHpel, int n,
                              const vtmme_mc_block* blocks, int16_t* dDst)
{
  if (!ctx) return VTMME_ERR_ARG;
  int    nTiles = 0, rc;
  size_t elems = 0;
  if ((rc = mc_check(ctx, "vtmme_mc_batch", comp, bitDepth, n, blocks, dDst, &nTiles, &elems)) != VTMME_OK) return rc;
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  return mc_launch(ctx, "vtmme_mc_batch", comp, bi, bitDepth, useAltHpel, n, blocks, nTiles, dDst);
}

extern "C" int vtmme_mc_host(vtmme_ctx* ctx, int comp, int bi, int bitDepth, int useAltHpel, int n,
                             const vtmme_mc_block* blocks, int16_t* dst)
{
  if (!ctx) return VTMME_ERR_ARG;
  int    nTiles = 0, rc;
  size_t elems = 0;
  if ((rc = mc_check(ctx, "vtmme_mc_host", comp, bitDepth, n, blocks, dst, &nTiles, &elems)) != VTMME_OK) return rc;
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // scratch below is shared with the other synchronous calls
  const size_t outBytes = align256(elems * 2);
  if ((rc = ensure_pinned(ctx, outBytes)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, outBytes)) != VTMME_OK) return rc;
  if ((rc = mc_launch(ctx, "vtmme_mc_host", comp, bi, bitDepth, useAltHpel, n, blocks, nTiles,
                      reinterpret_cast<int16_t*>(ctx->dJobBuf))) != VTMME_OK)
    return rc;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->hPinned, ctx->dJobBuf, elems * 2, cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(dst, ctx->hPinned, elems * 2);
  return VTMME_OK;
}

// ---- affine ME primitives (AffineGradientSearch's table entries) ---------------------------------------------------------
static int affine_dims_ok(int w, int h) { return w >= 4 && h >= 4 && w <= 128 && h <= 128; }

extern "C" int vtmme_affine_sobel_host(vtmme_ctx* ctx, int vertical, const int16_t* pred, int predStride, int w, int h, int32_t* deriv,
                                       int derivStride)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!pred || !deriv || !affine_dims_ok(w, h) || predStride < w || derivStride < w)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_affine_sobel_host", "bad argument (4 <= w, h <= 128)");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const size_t inBytes = align256((size_t) w * h * 2), outBytes = (size_t) w * h * 4;
  int rc;
  if ((rc = ensure_pinned(ctx, inBytes + outBytes)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, inBytes + outBytes)) != VTMME_OK) return rc;
  int16_t* hp = reinterpret_cast<int16_t*>(ctx->hPinned);
  for (int y = 0; y < h; y++) memcpy(hp + (size_t) y * w, pred + (ptrdiff_t) y * predStride, (size_t) w * 2);
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobBuf, ctx->hPinned, inBytes, cudaMemcpyHostToDevice, ctx->stream));
  int* dD = reinterpret_cast<int*>(ctx->dJobBuf + inBytes);
  VTMME_CUDA_CHECK(ctx, launch_affine_sobel(reinterpret_cast<const int16_t*>(ctx->dJobBuf), w, w, h, vertical, dD, ctx->stream));
  ctx->launches += 1;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->hPinned + inBytes, dD, outBytes, cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const int32_t* ho = reinterpret_cast<const int32_t*>(ctx->hPinned + inBytes);
  for (int y = 0; y < h; y++) memcpy(deriv + (ptrdiff_t) y * derivStride, ho + (size_t) y * w, (size_t) w * 4);
  return VTMME_OK;
}

extern "C" int vtmme_affine_equal_coeff_host(vtmme_ctx* ctx, const int16_t* residue, int residueStride, const int32_t* d0, const int32_t* d1,
                                             int derivStride, int w, int h, int sixParam, int64_t* coeff)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!residue || !d0 || !d1 || !coeff || !affine_dims_ok(w, h) || residueStride < w || derivStride < w)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_affine_equal_coeff_host", "bad argument (4 <= w, h <= 128)");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const size_t rBytes = align256((size_t) w * h * 2), dBytes = align256((size_t) w * h * 4), cBytes = 49 * 8;
  const size_t total = rBytes + 2 * dBytes + align256(cBytes);
  int rc;
  if ((rc = ensure_pinned(ctx, total)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, total)) != VTMME_OK) return rc;
  int16_t* hr = reinterpret_cast<int16_t*>(ctx->hPinned);
  int32_t* h0 = reinterpret_cast<int32_t*>(ctx->hPinned + rBytes);
  int32_t* h1 = reinterpret_cast<int32_t*>(ctx->hPinned + rBytes + dBytes);
  for (int y = 0; y < h; y++)
  {
    memcpy(hr + (size_t) y * w, residue + (ptrdiff_t) y * residueStride, (size_t) w * 2);
    memcpy(h0 + (size_t) y * w, d0 + (ptrdiff_t) y * derivStride, (size_t) w * 4);
    memcpy(h1 + (size_t) y * w, d1 + (ptrdiff_t) y * derivStride, (size_t) w * 4);
  }
  memcpy(ctx->hPinned + rBytes + 2 * dBytes, coeff, cBytes);   // the entries accumulate (the reference's caller zeroes them)
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobBuf, ctx->hPinned, rBytes + 2 * dBytes + cBytes, cudaMemcpyHostToDevice, ctx->stream));
  long long* dC = reinterpret_cast<long long*>(ctx->dJobBuf + rBytes + 2 * dBytes);
  VTMME_CUDA_CHECK(ctx, launch_affine_equal_coeff(reinterpret_cast<const int16_t*>(ctx->dJobBuf), w, reinterpret_cast<const int*>(ctx->dJobBuf + rBytes),
                                                  reinterpret_cast<const int*>(ctx->dJobBuf + rBytes + dBytes), w, w, h, sixParam, dC, ctx->stream));
  ctx->launches += 1;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->hPinned + rBytes + 2 * dBytes, dC, cBytes, cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(coeff, ctx->hPinned + rBytes + 2 * dBytes, cBytes);
  return VTMME_OK;
}

extern "C" int vtmme_affine_gradient_step(vtmme_ctx* ctx, int n, const vtmme_affine_block* blocks, int64_t* coeff)
{
  static const char* who = "vtmme_affine_gradient_step";
  if (!ctx) return VTMME_ERR_ARG;
  if (!blocks || !coeff || n <= 0 || n > 65536) return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  size_t sampleBytes = 0;
  for (int i = 0; i < n; i++)
  {
    const vtmme_affine_block& b = blocks[i];
    if (!b.org || !b.pred || !affine_dims_ok(b.w, b.h) || b.orgStride < b.w || b.predStride < b.w)
      return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "bad block (4 <= w, h <= 128)");
    sampleBytes += 2 * align256((size_t) b.w * b.h * 2);
  }
  const size_t descBytes = align256((size_t) n * sizeof(DevAffineBlock)), outBytes = align256((size_t) n * 49 * 8);
  const size_t total = descBytes + sampleBytes + outBytes;
  int rc;
  if ((rc = ensure_pinned(ctx, total)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, total)) != VTMME_OK) return rc;
  DevAffineBlock* hd = reinterpret_cast<DevAffineBlock*>(ctx->hPinned);
  size_t off = descBytes;
  for (int i = 0; i < n; i++)
  {
    const vtmme_affine_block& b = blocks[i];
    const size_t blk = align256((size_t) b.w * b.h * 2);
    int16_t* ho = reinterpret_cast<int16_t*>(ctx->hPinned + off);
    int16_t* hp = reinterpret_cast<int16_t*>(ctx->hPinned + off + blk);
    for (int y = 0; y < b.h; y++)
    {
      memcpy(ho + (size_t) y * b.w, b.org + (ptrdiff_t) y * b.orgStride, (size_t) b.w * 2);
      memcpy(hp + (size_t) y * b.w, b.pred + (ptrdiff_t) y * b.predStride, (size_t) b.w * 2);
    }
    hd[i].org  = reinterpret_cast<const int16_t*>(ctx->dJobBuf + off);
    hd[i].pred = reinterpret_cast<const int16_t*>(ctx->dJobBuf + off + blk);
    hd[i].orgStride = hd[i].predStride = b.w;
    hd[i].w = b.w;
    hd[i].h = b.h;
    hd[i].sixParam = b.sixParam != 0;
    hd[i].pad = 0;
    off += 2 * blk;
  }
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobBuf, ctx->hPinned, off, cudaMemcpyHostToDevice, ctx->stream));
  long long* dC = reinterpret_cast<long long*>(ctx->dJobBuf + off);
  VTMME_CUDA_CHECK(ctx, launch_affine_step(reinterpret_cast<const DevAffineBlock*>(ctx->dJobBuf), n, dC, ctx->stream));
  ctx->launches += 1;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->hPinned + off, dC, (size_t) n * 49 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(coeff, ctx->hPinned + off, (size_t) n * 49 * 8);
  return VTMME_OK;
}

// ---- symmetric-MVD search ------------------------------------------------------------------------------------------
static_assert(sizeof(vtmme_smvd_result) == sizeof(DevSmvdResult), "result layout");
extern "C" int vtmme_smvd_search(vtmme_ctx* ctx, int n, const vtmme_smvd* jobs, vtmme_smvd_result* results)
{
  static const char* who = "vtmme_smvd_search";
  if (!ctx) return VTMME_ERR_ARG;
  if (!jobs || !results || n <= 0 || n > 65536) return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  size_t orgBytes = 0;
  int    rc;
  for (int i = 0; i < n; i++)
  {
    const vtmme_smvd& j = jobs[i];
    const bool pow2w = j.w >= 8 && j.w <= 128 && (j.w & (j.w - 1)) == 0, pow2h = j.h >= 8 && j.h <= 128 && (j.h & (j.h - 1)) == 0;
    if (!pow2w || !pow2h || j.bitDepth < 8 || j.bitDepth > 10 || j.imv < 0 || j.imv > 3 || j.bcwIdx < 0 || j.bcwIdx > 4 ||
        j.maxCu < 1 || j.maxCu > 128 || (j.org && j.orgStride < j.w))
      return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "invalid job (w, h powers of two in 8..128, bit depth 8..10, imv 0..3, bcwIdx 0..4)");
    auto a = ctx->pics.find(j.refPicCur), b = ctx->pics.find(j.refPicTar);
    if (a == ctx->pics.end() || b == ctx->pics.end() || (!j.org && ctx->pics.find(j.curPic) == ctx->pics.end()))
      return vtmme_set_error(ctx, VTMME_ERR_NOPIC, who, "unknown picture id");
    const DevPic& pa = a->second;
    const DevPic& pb = b->second;
    if (pa.width != pb.width || pa.height != pb.height || pa.stride != pb.stride || pa.margin < 144 || pb.margin < 144)
      return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "the two reference pictures must have the same size");
    if (j.x < 0 || j.y < 0 || j.x + j.w > pa.width || j.y + j.h > pa.height)
      return vtmme_set_error(ctx, VTMME_ERR_RANGE, who, "PU outside the picture");
    if ((rc = wait_picture(ctx, j.refPicCur)) != VTMME_OK || (rc = wait_picture(ctx, j.refPicTar)) != VTMME_OK) return rc;
    if (!j.org && (rc = wait_picture(ctx, j.curPic)) != VTMME_OK) return rc;
    if (j.org) orgBytes += align256((size_t) j.w * j.h * 2);
  }
  const size_t jobBytes = align256((size_t) n * sizeof(DevSmvd)), resBytes = align256((size_t) n * sizeof(DevSmvdResult));
  const size_t total = jobBytes + resBytes + orgBytes;
  if ((rc = ensure_pinned(ctx, total)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, total)) != VTMME_OK) return rc;
  DevSmvd* hj  = reinterpret_cast<DevSmvd*>(ctx->hPinned);
  size_t   off = jobBytes + resBytes;
  for (int i = 0; i < n; i++)
  {
    const vtmme_smvd& j = jobs[i];
    const DevPic& pa = ctx->pics[j.refPicCur];
    const DevPic& pb = ctx->pics[j.refPicTar];
    DevSmvd& d = hj[i];
    if (j.org)
    {
      int16_t* ho = reinterpret_cast<int16_t*>(ctx->hPinned + off);
      for (int y = 0; y < j.h; y++) memcpy(ho + (size_t) y * j.w, j.org + (ptrdiff_t) y * j.orgStride, (size_t) j.w * 2);
      d.org       = reinterpret_cast<const int16_t*>(ctx->dJobBuf + off);
      d.orgStride = j.w;
      off += align256((size_t) j.w * j.h * 2);
    }
    else
    {
      const DevPic& cp = ctx->pics[j.curPic];
      d.org       = cp.origin + (ptrdiff_t) j.y * cp.stride + j.x;
      d.orgStride = cp.stride;
    }
    d.refCur = pa.origin;
    d.refTar = pb.origin;
    d.refStride = pa.stride;
    d.x = j.x; d.y = j.y; d.w = j.w; d.h = j.h;
    d.picW = pa.width; d.picH = pa.height; d.maxCu = j.maxCu; d.bd = j.bitDepth; d.imv = j.imv;
    d.curPredX = j.curPredX; d.curPredY = j.curPredY; d.tarPredX = j.tarPredX; d.tarPredY = j.tarPredY;
    d.curMvX = j.curMvX; d.curMvY = j.curMvY; d.tarMvX = j.tarMvX; d.tarMvY = j.tarMvY;
    d.clipBiPred = j.clipBiPred; d.useHad = j.useHad; d.bcwIdx = j.bcwIdx;
    d.lambda = j.lambdaMotion;
    d.cost   = j.cost;
  }
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobBuf, ctx->hPinned, total, cudaMemcpyHostToDevice, ctx->stream));
  DevSmvdResult* dRes = reinterpret_cast<DevSmvdResult*>(ctx->dJobBuf + jobBytes);
  VTMME_CUDA_CHECK(ctx, launch_smvd_search(reinterpret_cast<const DevSmvd*>(ctx->dJobBuf), dRes, n, ctx->stream));
  ctx->launches += 1;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->hPinned + jobBytes, dRes, (size_t) n * sizeof(DevSmvdResult), cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(results, ctx->hPinned + jobBytes, (size_t) n * sizeof(vtmme_smvd_result));
  return VTMME_OK;
}

// Mv clipping of clipMvInPic (CommonLib/Mv.cpp:53-71), 1/16 sample
static void clip_mv_in_pic(int& mx, int& my, int x, int y, int picW, int picH, int maxCu)
{
  const int horMax = (picW + 8 - x - 1) << 4, horMin = (-maxCu - 8 - x + 1) * 16;
  const int verMax = (picH + 8 - y - 1) << 4, verMin = (-maxCu - 8 - y + 1) * 16;
  mx = mx > horMax ? horMax : (mx < horMin ? horMin : mx);
  my = my > verMax ? verMax : (my < verMin ? verMin : my);
}

// The prediction of one list after DMVR: xFinalPaddedMCForDMVR (CommonLib/InterPrediction.cpp:1845-1917).  xPrefetch (:1664-1708)
// copied the (w + taps - 1) x (h + taps - 1) window at the integer part of clip(mergeMv - (taps/2 - 1) samples), xPad
// (:1710-1730) replicated its border, and xPredInterBlk (bi = true) filters out of that buffer at the refined MV: the same
// tile routine as vtmme_mc_batch reading the window with clamped coordinates.
extern "C" int vtmme_dmvr_final_mc(vtmme_ctx* ctx, int comp, int refPic, int bitDepth, int maxCu, int n,
                                   const vtmme_dmvr_block* blocks, int16_t* dst)
{
  static const char* who = "vtmme_dmvr_final_mc";
  if (!ctx) return VTMME_ERR_ARG;
  if (!blocks || !dst || n <= 0 || n > (1 << 20) || (comp != 0 && comp != 1) || bitDepth < 8 || bitDepth > 10 || maxCu < 1 || maxCu > 128)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  auto it = ctx->pics.find(refPic);
  if (it == ctx->pics.end()) return vtmme_set_error(ctx, VTMME_ERR_NOPIC, who, "unknown picture id");
  int rc;
  if ((rc = wait_picture(ctx, refPic)) != VTMME_OK) return rc;
  const DevPic rp = it->second;
  const int cs = comp ? 1 : 0, taps = comp ? 4 : 8, before = taps / 2 - 1, sh = 4 + cs;
  const int picW = rp.width << cs, picH = rp.height << cs;   // the luma picture the MV clip refers to
  size_t elems = 0;
  for (int i = 0; i < n; i++)
  {
    const vtmme_dmvr_block& b = blocks[i];
    if ((b.w != 8 && b.w != 16) || (b.h != 8 && b.h != 16) || b.x < 0 || b.y < 0 || b.x + b.w > picW || b.y + b.h > picH)
      return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "sub-blocks are 8 or 16 luma samples wide / high and lie inside the picture");
    elems += (size_t) (b.w >> cs) * (b.h >> cs);
  }
  if ((rc = mc_tiles_begin(ctx, n)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, elems * 2)) != VTMME_OK) return rc;
  if ((rc = ensure_pinned(ctx, elems * 2)) != VTMME_OK) return rc;
  int16_t* dDst = reinterpret_cast<int16_t*>(ctx->dJobBuf);
  McTile*  ht   = reinterpret_cast<McTile*>(ctx->hMcTiles);
  size_t   off  = 0;
  for (int i = 0; i < n; i++)
  {
    const vtmme_dmvr_block& b = blocks[i];
    const int w = b.w >> cs, h = b.h >> cs;
    int cx = b.mvL0x - (before << sh), cy = b.mvL0y - (before << sh), fx = b.mvL1x, fy = b.mvL1y;   // merge MV -> window, refined MV -> phase
    clip_mv_in_pic(cx, cy, b.x, b.y, picW, picH, maxCu);
    clip_mv_in_pic(fx, fy, b.x, b.y, picW, picH, maxCu);
    const int wx0 = (b.x >> cs) + (cx >> sh), wy0 = (b.y >> cs) + (cy >> sh);   // window origin in the plane
    if (wx0 < -rp.margin || wy0 < -rp.margin || wx0 + w + taps - 1 > rp.width + rp.margin || wy0 + h + taps - 1 > rp.height + rp.margin)
      return vtmme_set_error(ctx, VTMME_ERR_RANGE, who, "prefetch window leaves the padded reference plane");
    const int bx = before + ((b.mvL1x >> sh) - (b.mvL0x >> sh)), by = before + ((b.mvL1y >> sh) - (b.mvL0y >> sh));
    if (bx < -100 || bx > 100 || by < -100 || by > 100)
      return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "refined MV too far from the merge MV (DMVR moves a block by at most 2 samples)");
    McTile& t = ht[i];
    t.src = rp.origin + (ptrdiff_t) wy0 * rp.stride + wx0;
    t.dst = dDst + off;
    t.srcStride = rp.stride;
    t.dstStride = w;
    t.tw = (uint8_t) w;
    t.th = (uint8_t) h;
    t.xFrac = (uint8_t) (fx & ((1 << sh) - 1));
    t.yFrac = (uint8_t) (fy & ((1 << sh) - 1));
    t.q4Hor = t.q4Ver = 0;
    t.winMaxX = (uint8_t) (w + taps - 2);
    t.winMaxY = (uint8_t) (h + taps - 2);
    t.winX = (int8_t) bx;
    t.winY = (int8_t) by;
    off += (size_t) w * h;
  }
  if ((rc = mc_tiles_launch(ctx, comp, n, 1, bitDepth, 0)) != VTMME_OK) return rc;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->hPinned, dDst, elems * 2, cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(dst, ctx->hPinned, elems * 2);
  return VTMME_OK;
}

extern "C" int vtmme_add_avg(vtmme_ctx* ctx, const int16_t* dSrc0, const int16_t* dSrc1, int16_t* dDst, int64_t count,
                             int bitDepth)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!dSrc0 || !dSrc1 || !dDst || count <= 0 || bitDepth < 8 || bitDepth > 10)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_add_avg", "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  VTMME_CUDA_CHECK(ctx, launch_add_avg(dSrc0, dSrc1, dDst, count, bitDepth, ctx->stream));
  ctx->launches += 1;
  return VTMME_OK;
}

extern "C" int vtmme_remove_high_freq(vtmme_ctx* ctx, int16_t* dOrg, const int16_t* dPred, int64_t count, int clip,
                                      int bitDepth)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!dOrg || !dPred || count <= 0 || bitDepth < 8 || bitDepth > 10)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_remove_high_freq", "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  VTMME_CUDA_CHECK(ctx, launch_remove_high_freq(dOrg, dPred, count, clip, bitDepth, ctx->stream));
  ctx->launches += 1;
  return VTMME_OK;
}

extern "C" int vtmme_add_weighted_avg(vtmme_ctx* ctx, const int16_t* dSrc0, const int16_t* dSrc1, int16_t* dDst, int64_t count,
                                      int bitDepth, int bcwIdx)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!dSrc0 || !dSrc1 || !dDst || count <= 0 || bitDepth < 8 || bitDepth > 10 || bcwIdx < 0 || bcwIdx > 4)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_add_weighted_avg", "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  VTMME_CUDA_CHECK(ctx, launch_add_weighted_avg(dSrc0, dSrc1, dDst, count, bitDepth, bcwIdx, ctx->stream));
  ctx->launches += 1;
  return VTMME_OK;
}

extern "C" int vtmme_remove_weight_high_freq(vtmme_ctx* ctx, int16_t* dOrg, const int16_t* dPred, int64_t count, int clip,
                                             int bitDepth, int bcwWeight)
{
  if (!ctx) return VTMME_ERR_ARG;
  if (!dOrg || !dPred || count <= 0 || bitDepth < 8 || bitDepth > 10 || bcwWeight == 0 || bcwWeight < -8 || bcwWeight > 16)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, "vtmme_remove_weight_high_freq", "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  VTMME_CUDA_CHECK(ctx, launch_remove_weight_high_freq(dOrg, dPred, count, clip, bitDepth, bcwWeight, ctx->stream));
  ctx->launches += 1;
  return VTMME_OK;
}

// ---- candidate distortion (template cost / seeds) -------------------------------------------------------------------
static_assert(sizeof(vtmme_dmvr_block) == sizeof(DevDmvrBlock) && sizeof(vtmme_dmvr_result) == sizeof(DevDmvrResult),
              "DMVR layouts must match");

extern "C" int vtmme_dmvr_refine(vtmme_ctx* ctx, int refPic0, int refPic1, int bitDepth, int maxCu, int n,
                                 const vtmme_dmvr_block* blocks, vtmme_dmvr_result* results)
{
  if (!ctx) return VTMME_ERR_ARG;
  const char* who = "vtmme_dmvr_refine";
  if (!blocks || !results || n <= 0 || n > (1 << 20) || bitDepth < 8 || bitDepth > 10 || maxCu < 8 || maxCu > 128)
    return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "bad argument (at most 2^20 sub-blocks per call, bitDepth 8..10, maxCu 8..128)");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  if (ctx->pics.find(refPic0) == ctx->pics.end() || ctx->pics.find(refPic1) == ctx->pics.end())
    return vtmme_set_error(ctx, VTMME_ERR_NOPIC, who, "unknown picture id");
  const DevPic r0 = ctx->pics[refPic0], r1 = ctx->pics[refPic1];
  if (r0.width != r1.width || r0.height != r1.height) return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "the two reference pictures differ in size");
  for (int i = 0; i < n; i++)
  {
    const vtmme_dmvr_block& b = blocks[i];
    if ((b.w != 8 && b.w != 16) || (b.h != 8 && b.h != 16) || b.x < 0 || b.y < 0 || b.x + b.w > r0.width || b.y + b.h > r0.height)
      return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "bad sub-block (w,h in {8,16}, inside the picture)");
  }
  // the MV clip keeps every fetched sample within maxCu + 8 + 16 + 7 samples of the picture: inside the device margin
  int rc;
  if ((rc = wait_picture(ctx, refPic0)) != VTMME_OK || (rc = wait_picture(ctx, refPic1)) != VTMME_OK) return rc;
  const size_t inBytes = align256((size_t) n * sizeof(DevDmvrBlock)), total = inBytes + align256((size_t) n * sizeof(DevDmvrResult));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // the staging blocks are shared with the other synchronous calls
  if ((rc = ensure_pinned(ctx, total)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, total)) != VTMME_OK) return rc;
  memcpy(ctx->hPinned, blocks, (size_t) n * sizeof(DevDmvrBlock));
  DevDmvrResult* dRes = reinterpret_cast<DevDmvrResult*>(ctx->dJobBuf + inBytes);
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobBuf, ctx->hPinned, (size_t) n * sizeof(DevDmvrBlock), cudaMemcpyHostToDevice, ctx->stream));
  VTMME_CUDA_CHECK(ctx, launch_dmvr_refine(r0, r1, reinterpret_cast<const DevDmvrBlock*>(ctx->dJobBuf), n, bitDepth, maxCu, dRes, ctx->stream));
  ctx->launches += 1;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->hPinned + inBytes, dRes, (size_t) n * sizeof(DevDmvrResult), cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(results, ctx->hPinned + inBytes, (size_t) n * sizeof(DevDmvrResult));
  return VTMME_OK;
}

extern "C" int vtmme_cand_sad(vtmme_ctx* ctx, int bitDepth, int useAltHpel, int nJobs, const vtmme_cand_job* jobs, uint64_t* out)
{
  if (!ctx) return VTMME_ERR_ARG;
  const char* who = "vtmme_cand_sad";
  if (!jobs || !out || nJobs <= 0 || bitDepth < 8 || bitDepth > 10) return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "bad argument");
  VTMME_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  size_t    orgBytes = 0;
  long long nTiles = 0, nCand = 0;
  for (int i = 0; i < nJobs; i++)
  {
    const vtmme_cand_job& j = jobs[i];
    if (j.w < 1 || j.h < 1 || j.w > 128 || j.h > 128 || j.nCand < 1 || j.nCand > 64 || !j.mv || j.subShift < 0 || j.subShift > 4)
      return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "bad job (w,h in [1,128], nCand in [1,64], subShift in [0,4])");
    if (ctx->pics.find(j.refPic) == ctx->pics.end() || (!j.org && ctx->pics.find(j.curPic) == ctx->pics.end()))
      return vtmme_set_error(ctx, VTMME_ERR_NOPIC, who, "unknown picture id");
    if (j.org) orgBytes += align256((size_t) j.w * j.h * 2);
    nTiles += (long long) j.nCand * ((j.w + 15) / 16) * ((j.h + 15) / 16);
    nCand += j.nCand;
  }
  if (nTiles > (1 << 24)) return vtmme_set_error(ctx, VTMME_ERR_ARG, who, "too many candidates in one call");
  // one pinned block / one device block: [tiles][patterns][sums]
  const size_t offOrg = align256((size_t) nTiles * sizeof(McSadTile)), offOut = align256(offOrg + orgBytes);
  const size_t total = offOut + align256((size_t) nCand * 8);
  int rc;
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // the staging blocks are shared with the other synchronous calls
  if ((rc = ensure_pinned(ctx, total)) != VTMME_OK) return rc;
  if ((rc = ensure(ctx, ctx->dJobBuf, ctx->jobBufCap, total)) != VTMME_OK) return rc;
  McSadTile* ht = reinterpret_cast<McSadTile*>(ctx->hPinned);
  size_t     orgCur = offOrg;
  int        k = 0, slot = 0;
  for (int i = 0; i < nJobs; i++)
  {
    const vtmme_cand_job& j = jobs[i];
    const DevPic& rp = ctx->pics[j.refPic];
    int wrc;
    if ((wrc = wait_picture(ctx, j.refPic)) != VTMME_OK) return wrc;
    const int16_t* dOrg;
    int            orgStride;
    if (j.org)
    {
      int16_t* ho = reinterpret_cast<int16_t*>(ctx->hPinned + orgCur);
      for (int y = 0; y < j.h; y++) memcpy(ho + (size_t) y * j.w, j.org + (ptrdiff_t) y * j.orgStride, (size_t) j.w * 2);
      dOrg      = reinterpret_cast<const int16_t*>(ctx->dJobBuf + orgCur);
      orgStride = j.w;
      orgCur += align256((size_t) j.w * j.h * 2);
    }
    else
    {
      const DevPic& cp = ctx->pics[j.curPic];
      if ((wrc = wait_picture(ctx, j.curPic)) != VTMME_OK) return wrc;
      if (j.x < 0 || j.y < 0 || j.x + j.w > cp.width || j.y + j.h > cp.height)
        return vtmme_set_error(ctx, VTMME_ERR_RANGE, who, "PU outside the original picture");
      dOrg      = cp.origin + (ptrdiff_t) j.y * cp.stride + j.x;
      orgStride = cp.stride;
    }
    for (int c = 0; c < j.nCand; c++, slot++)
    {
      const int mvX = j.mv[2 * c], mvY = j.mv[2 * c + 1];
      const int ix = mvX >> 4, iy = mvY >> 4, xFrac = mvX & 15, yFrac = mvY & 15;
      if (j.x + ix - 3 < -rp.margin || j.x + j.w + ix + 4 > rp.width + rp.margin || j.y + iy - 3 < -rp.margin ||
          j.y + j.h + iy + 4 > rp.height + rp.margin)
        return vtmme_set_error(ctx, VTMME_ERR_RANGE, who, "candidate leaves the padded reference plane (clip the MV first)");
      const int  hHor  = yFrac ? j.h + 7 : j.h;
      const bool q4Hor = j.w == 4 && (hHor == 4 || hHor == 11), q4Ver = j.w == 4 && j.h == 4;
      const int16_t* src = rp.origin + (ptrdiff_t) (j.y + iy) * rp.stride + j.x + ix;
      for (int ty = 0; ty < j.h; ty += 16)
        for (int tx = 0; tx < j.w; tx += 16)
        {
          McSadTile& t = ht[k++];
          t.mc.src = src + (ptrdiff_t) ty * rp.stride + tx;
          t.mc.dst = nullptr;
          t.mc.srcStride = rp.stride;
          t.mc.dstStride = 0;
          t.mc.tw = (uint8_t) (j.w - tx < 16 ? j.w - tx : 16);
          t.mc.th = (uint8_t) (j.h - ty < 16 ? j.h - ty : 16);
          t.mc.xFrac = (uint8_t) xFrac;
          t.mc.yFrac = (uint8_t) yFrac;
          t.mc.q4Hor = q4Hor;
          t.mc.q4Ver = q4Ver;
          t.mc.winMaxX = t.mc.winMaxY = 0;
          t.mc.winX = t.mc.winY = 0;
          t.org = dOrg + (ptrdiff_t) ty * orgStride + tx;
          t.orgStride = orgStride;
          t.outIdx = slot;
          t.subShift = j.subShift;
          t.pad = 0;
        }
    }
  }
  unsigned long long* dOut = reinterpret_cast<unsigned long long*>(ctx->dJobBuf + offOut);
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->dJobBuf, ctx->hPinned, offOrg + orgBytes, cudaMemcpyHostToDevice, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaMemsetAsync(dOut, 0, (size_t) nCand * 8, ctx->stream));
  VTMME_CUDA_CHECK(ctx, launch_mc_sad(reinterpret_cast<const McSadTile*>(ctx->dJobBuf), (int) nTiles, bitDepth, useAltHpel, dOut,
                                      ctx->stream));
  ctx->launches += 1;
  VTMME_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->hPinned + offOut, dOut, (size_t) nCand * 8, cudaMemcpyDeviceToHost, ctx->stream));
  VTMME_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->hPinned + offOut, (size_t) nCand * 8);
  return VTMME_OK;
}
