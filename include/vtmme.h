/* libvtmme — B200-native (sm_100a CUDA) motion search for VTM 9.3, C ABI.
 *
 * This is the drop-in boundary for ONE path of the reference encoder: the inter-prediction motion
 * search (integer full search + half/quarter-pel refinement) and the distortion / interpolation
 * kernels it bottoms out in.  Plain C, POD arguments, no exceptions; every entry point returns
 * VTMME_OK (0) or a negative error code and never falls back to a CPU implementation — when the CUDA
 * path cannot run, the call fails and the caller (VTM: THROW) must stop.
 *
 * Paths below are relative to the reference's source/Lib directory.
 *
 *   reference interface replaced                                   entry point here
 *   ------------------------------------------------------------  ---------------------------------
 *   RdCost::m_afpDistortFunc[DF_SAD*]  (RdCost.h:113,               vtmme_dist_batch (kind 0) / vtmme_dist_host
 *     RdCost.cpp:125-208; x86/RdCostX86.h:2307-2321)
 *   RdCost::m_afpDistortFunc[DF_HAD*]  (x86/RdCostX86.h:2323-2330)  vtmme_dist_batch (kind 1) / vtmme_dist_host
 *   InterpolationFilter::m_filterHor/m_filterVer/m_filterCopy       vtmme_interp_batch / vtmme_interp_host
 *     (InterpolationFilter.h:96-98, InterpolationFilter.cpp:749-895)
 *   InterSearch::xPatternSearch + xPatternSearchFracDIF             vtmme_search          (per-call jobs)
 *     (EncoderLib/InterSearch.cpp:3566-3608, 4284-4339)             vtmme_search_frames   (batched, per CTU tree)
 *   InterSearch::xPatternSearchIntRefine (:4172-4282)               vtmme_search with fracMode 2 + vtmme_amvr
 *   InterSearch::xTZSearch (:3640-3974, FastSearch=1/3)             vtmme_search with vtmme_tz
 *   InterSearch::xTZSearchSelective (:3979-4170, FastSearch=2)      vtmme_search with vtmme_tz.selective
 *   InterPrediction::xPredInterBlk (CommonLib/InterPrediction.cpp   vtmme_mc_batch / vtmme_mc_host
 *     :660-830), AreaBuf::addAvg (Buffer.cpp:467-507),              vtmme_add_avg
 *     AreaBuf::removeHighFreq (Buffer.h:474-517)                    vtmme_remove_high_freq
 *   EncTemporalFilter::motionEstimation / applyMotion (EncoderLib/  vtmme_mctf_me / vtmme_mctf_apply_motion
 *     EncTemporalFilter.cpp:448-466, 470-552)
 *   InterPrediction::xProcessDMVR, the search (CommonLib/           vtmme_dmvr_refine
 *     InterPrediction.cpp:2098-2154)
 *   distortion of InterSearch::xGetTemplateCost and of the ME      vtmme_cand_sad
 *     seeds (EncoderLib/InterSearch.cpp:3235-3270, 3388-3426)
 *   Picture::getRecoBuf / getOrigBuf planes handed to ME            vtmme_upload_picture / vtmme_release_picture
 *     (Picture.cpp:322-329, extendPicBorder :1050-1110)
 *
 * Threading: one context per encoder thread; calls on one context are serialised by the caller
 * (VTM is single-threaded).  All host-pointer entry points are synchronous on return.
 */
#ifndef VTMME_H
#define VTMME_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VTMME_OK 0
#define VTMME_ERR_CUDA (-1)    /* a CUDA runtime call or kernel failed; see vtmme_last_error */
#define VTMME_ERR_ARG (-2)     /* invalid argument */
#define VTMME_ERR_NOMEM (-3)   /* device allocation failed */
#define VTMME_ERR_NOPIC (-4)   /* unknown picture id */
#define VTMME_ERR_RANGE (-5)   /* a window / predictor spread exceeds what the context was sized for */

typedef struct vtmme_ctx vtmme_ctx;

/* ---- context ------------------------------------------------------------------------------- */
int         vtmme_create(int device, vtmme_ctx** ctx);
void        vtmme_destroy(vtmme_ctx* ctx);
const char* vtmme_last_error(const vtmme_ctx* ctx);
/* Run on the caller's CUDA stream (a cudaStream_t, e.g. a torch.cuda.Stream's .cuda_stream); NULL = the library's
 * own non-blocking stream.  (The legacy default stream's handle is NULL too: pass cudaStreamLegacy for it.) */
int         vtmme_set_stream(vtmme_ctx* ctx, void* cudaStream);
int         vtmme_synchronize(vtmme_ctx* ctx);
/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t    vtmme_launch_count(const vtmme_ctx* ctx);

/* ---- pictures -------------------------------------------------------------------------------
 * A picture is one int16 luma plane (Pel, TypeDef.h:259).  `origin` points at sample (0,0).
 * withBorder != 0: the caller's plane already carries `margin` valid samples on every side (what
 *   Picture::extendPicBorder produced, Picture.cpp:1050-1110) and they are copied as they are.
 * withBorder == 0: only the width x height area is read and the library replicates the border on
 *   the device (same rule as extendPicBorder: edge samples repeated).
 * The device copy always has a margin of at least VTMME_MIN_MARGIN samples.                     */
#define VTMME_MIN_MARGIN 192
int vtmme_upload_picture(vtmme_ctx* ctx, int picId, const int16_t* origin, int stride, int width, int height,
                         int margin, int withBorder);
/* Asynchronous flavour for pipelining: the copy and the border extension run on an internal copy stream and overlap
 * with searches on other pictures; any later search / job that names picId waits for it on the device.  `origin` must
 * be page-locked and stay valid until vtmme_synchronize() (or a synchronous call that used the picture) returned. */
int vtmme_upload_picture_async(vtmme_ctx* ctx, int picId, const int16_t* origin, int stride, int width, int height,
                               int margin, int withBorder);
/* Same, source plane already in device memory (copied device-to-device on the context stream). */
int vtmme_upload_picture_device(vtmme_ctx* ctx, int picId, const int16_t* dOrigin, int stride, int width, int height,
                                int margin, int withBorder);
int vtmme_release_picture(vtmme_ctx* ctx, int picId);

/* ---- per-call motion search ------------------------------------------------------------------
 * One job = one InterSearch::xMotionEstimation call's integer search + refinement:
 * xPatternSearch over searchRange (InterSearch.cpp:3566-3608), then the body of
 * xPatternSearchFracDIF (:4296-4338) for quarter / half-pel CUs, or xPatternSearchIntRefine
 * (:4172-4282) for integer / 4-pel AMVR CUs.  Field meaning follows IntTZSearchStruct / DistParam. */
/* State of an integer-pel / 4-pel AMVR call (cu.imv = IMV_FPEL / IMV_4PEL): what
 * InterSearch::xPatternSearchIntRefine (InterSearch.cpp:4172-4282) receives besides the search result.
 * MVs in 1/16 sample (MV_PRECISION_INTERNAL). */
typedef struct
{
  int32_t  imv;            /* cu.imv: 1 IMV_FPEL, 2 IMV_4PEL                                           */
  int32_t  numCand;        /* amvpInfo.numCand, 1 or 2                                                 */
  int32_t  candX[2], candY[2]; /* amvpInfo.mvCand[]                                                    */
  int32_t  mvpIdx;         /* riMVPIdx on entry (mvCand[mvpIdx] == rcMvPred)                           */
  uint32_t mvpIdxBits[2];  /* m_auiMVPIdxCost[i][AMVP_MAX_NUM_CANDS]                                   */
  uint32_t bits;           /* ruiBits on entry                                                         */
  int32_t  picW, picH;     /* pps.getPicWidth/HeightInLumaSamples (clipMvInPic, Mv.cpp:53-71)          */
  int32_t  maxCuW, maxCuH; /* sps.getMaxCUWidth/Height, <= 128                                         */
  double   fWeight;        /* xGetMEDistortionWeight (InterSearch.cpp:7666-7676)                       */
} vtmme_amvr;

/* State of a TZ search call, InterSearch::xTZSearch (InterSearch.cpp:3640-3974): when a job carries one, its
 * integer search is the TZ search of FastSearch=1 (MESEARCH_DIAMOND: extended 0, fast 0), FastSearch=3
 * (MESEARCH_DIAMOND_ENHANCED: extended 1), of the cached-MV re-search (:3445, fast 1) or the selective search of
 * FastSearch=2 (xTZSearchSelective :3979-4170, selective 1) instead of the full
 * search; the job's srLeft..srBottom are ignored (xTZSearch sets its own window around the best start point).
 * No hash ME, MCTS or composite reference; subShift as RdCost::setDistParam gives it for subShiftMode 0, 1 or 2. */
typedef struct
{
  int32_t startX, startY;       /* rcMv on entry (the AMVP predictor, or the cached integer MV), 1/16 sample      */
  int32_t hasInt2Nx2N;          /* pIntegerMv2Nx2NPred != NULL                                                    */
  int32_t int2Nx2NX, int2Nx2NY; /*   its value, integer pel                                                       */
  int32_t nSeeds;               /* m_uniMvListSize, 0..15                                                         */
  int32_t seedX[16], seedY[16]; /* uniMvs[list][ref] of the history entries, newest first (duplicates allowed,   */
                                /*   the search skips them like :3738-3748), 1/16 sample                          */
  int32_t searchRange;          /* m_iSearchRange                                                                 */
  int32_t extended, fast;       /* bExtendedSettings, bFastSettings                                               */
  int32_t firstSearchStop;      /* EncCfg::getFastMEAssumingSmootherMVEnabled                                     */
  int32_t picW, picH;           /* clipMvInPic / xClipMv rectangle = size of refPic                               */
  int32_t maxCu;                /* sps.getMaxCUWidth() == getMaxCUHeight(), <= 128                                */
  int32_t selective;            /* 1: xTZSearchSelective (:3979-4170, FastSearch=2 / MESEARCH_SELECTIVE) instead of */
                                /*   xTZSearch; extended, fast, firstSearchStop unused; searchRange <= 128         */
  int32_t stagedSad;            /* 1: cStruct.subShiftMode == 1 (what MESEARCH_SELECTIVE sets without               */
                                /*   RestrictMESampling, :3438-3445): every probe is xTZSearchHelp's staged SAD     */
                                /*   (:340-391) and the job's subShift is setDistParam's mode-1 value               */
                                /*   (RdCost.cpp:291-309: 4 for h > 32, 3 for h 32, 2 for h 16, else 1)             */
} vtmme_tz;

typedef struct
{
  int32_t        curPic;       /* picture holding the original block; ignored when org != NULL        */
  int32_t        refPic;       /* reference picture (recon plane with border)                          */
  int32_t        x, y, w, h;   /* PU luma rectangle; w,h in {4,8,16,32,64,128}                         */
  const int16_t* org;          /* optional HOST pointer to the pattern (bi-pred: 2*org - otherPred,    */
  int32_t        orgStride;    /*   InterSearch.cpp:3317-3328); NULL = read curPic at (x,y)            */
  int32_t        srLeft, srRight, srTop, srBottom; /* cStruct.searchRange (integer pel, inclusive)     */
  int32_t        predQx, predQy;                   /* RdCost::setPredictor value, quarter-pel          */
  int32_t        imvShift;     /* 0 qpel, 1 hpel, 2 fpel, 4 4pel (InterSearch.cpp:3344)                */
  int32_t        subShift;     /* DistParam::subShift of the integer search (RdCost.cpp:289-323)       */
  int32_t        bitDepth;     /* <= 10 (larger falls outside the SIMD tables too: RdCostX86.h:213)    */
  int32_t        useHad;       /* HadamardME && !DisableSATDForRD                                      */
  int32_t        useAltHpel;   /* cStruct.useAltHpelIf                                                 */
  int32_t        fracMode;     /* 0: integer only; 1: xPatternSearchFracDIF body; 2: the search is     */
                               /*   followed by xPatternSearchIntRefine on `amvr` (:3486-3489)         */
  double         lambdaMotion; /* RdCost::m_motionLambda                                               */
  const vtmme_amvr* amvr;      /* HOST pointer, fracMode 2 only (else ignored, may be NULL)            */
  const vtmme_tz*   tz;        /* HOST pointer or NULL; non-NULL: xTZSearch instead of xPatternSearch  */
                               /*   (all jobs of one call must agree)                                  */
} vtmme_job;

typedef struct
{
  int32_t  mvX, mvY;      /* best integer MV (xPatternSearch rcMv)                                     */
  uint64_t intSad;        /* ruiSAD: best cost minus its MV cost (InterSearch.cpp:3606)                */
  int32_t  halfX, halfY;  /* rcMvHalf                                                                  */
  int32_t  qterX, qterY;  /* rcMvQter                                                                  */
  uint64_t fracCost;      /* ruiCost after xPatternSearchFracDIF                                       */
  /* fracMode 2 only — outputs of xPatternSearchIntRefine: */
  int32_t  amvrMvX, amvrMvY; /* rcMv, 1/16 sample                                                      */
  int32_t  mvpIdx;        /* riMVPIdx (rcMvPred = mvCand[mvpIdx])                                      */
  uint32_t bits;          /* ruiBits                                                                   */
  uint64_t cost;          /* ruiCost                                                                   */
} vtmme_result;

int vtmme_search(vtmme_ctx* ctx, const vtmme_job* jobs, int n, vtmme_result* results);

/* ---- batched per-CTU search (standalone ME, BASELINE config 4) ---------------------------------
 * For every picture pair, every grid-aligned square CU of size 8,16,32,64,128 fully inside the
 * picture is searched: window = xSetSearchRange(pred, searchRange) (InterSearch.cpp:3496-3535),
 * integer full search, then half + quarter-pel refinement with SATD.  SADs of nested CUs are
 * computed once per 8x8 block and displacement and summed up the quad-tree.
 *
 * CU order in pred / result arrays: level-major, level l = CUs of size 8<<l (l = 0..4), each level in
 * raster order over floor(width/size) x floor(height/size) CUs.                                   */
typedef struct
{
  int32_t searchRange;   /* SearchRange (integer pel)                                               */
  int32_t bitDepth;      /* internal bit depth (10)                                                 */
  int32_t ctuSize;       /* sps.getMaxCUWidth() used by the MV clip (128)                           */
  int32_t imvShift;      /* 0                                                                       */
  int32_t useHad;        /* 1: SATD in the fractional stage                                         */
  int32_t fracMode;      /* 0 integer only, 1 half + quarter                                        */
  int32_t predSpread;    /* max |pred_a - pred_b| (integer pel, per component) among CUs of one CTU  */
  int32_t subShiftMode;  /* 0: every row; 2 (FEN=1/3): CUs with H > 8 and W <= 64 use even rows only, x2 (RdCost.cpp:310-316); */
                         /*   1 (fastSearch 2 only): xTZSearchHelp's staged SAD (InterSearch.cpp:340-391), what FastSearch=2 */
                         /*   selects without RestrictMESampling                                                          */
  double  lambdaMotion;
  int32_t fastSearch;    /* integer search, like --FastSearch: 0 full search (xPatternSearch); 1 TZ search (xTZSearch,  */
                         /*   MESEARCH_DIAMOND) started at the predictor; 3 enhanced TZ search (MESEARCH_DIAMOND_ENHANCED); */
                         /*   2 selective TZ search (xTZSearchSelective, MESEARCH_SELECTIVE; searchRange <= 128)            */
  int32_t tzFirstSearchStop; /* EncCfg::getFastMEAssumingSmootherMVEnabled (VTM default 1); fastSearch 1 / 3 only */
} vtmme_frame_params;

typedef struct
{
  int16_t  mvQx, mvQy;   /* final MV, quarter-pel: (int<<2) + (half<<1) + qter (InterSearch.cpp:3478-3480) */
  int16_t  intX, intY;   /* best integer MV                                                         */
  uint32_t intSad;       /* ruiSAD of xPatternSearch                                                */
  uint32_t fracCost;     /* ruiCost after the fractional stage (= intSad + mv cost when fracMode 0)  */
} vtmme_cu_result;

/* Number of CUs per picture and the offset of each of the 5 levels in the CU order. */
int vtmme_frame_cu_count(int width, int height, int32_t levelOffset[6]);

/* Host buffers: predQ = nPairs * nCU * 2 int16 (quarter-pel x,y) or NULL for zero predictors;
 * results = nPairs * nCU entries.  Synchronous. */
int vtmme_search_frames(vtmme_ctx* ctx, int nPairs, const int32_t* curPics, const int32_t* refPics,
                        const vtmme_frame_params* params, const int16_t* predQ, vtmme_cu_result* results);
/* Device buffers, asynchronous on the context stream (vtmme_synchronize to wait). */
int vtmme_search_frames_device(vtmme_ctx* ctx, int nPairs, const int32_t* curPics, const int32_t* refPics,
                               const vtmme_frame_params* params, const int16_t* dPredQ, vtmme_cu_result* dResults);

/* ---- table-level kernels (dispatch-table flavour; BASELINE config 5) ----------------------------
 * n independent block pairs.  Block i: org at org + i*orgBlockStride (row stride orgStride), cur
 * likewise; all blocks w x h.  kind 0: SAD with row sub-sampling subShift (RdCost.cpp:493-528);
 * kind 1: SATD with the reference tiling (RdCost.cpp:2819-2934).  Pointers are DEVICE pointers;
 * out = n uint64 (Distortion, TypeDef.h:270). */
int vtmme_dist_batch(vtmme_ctx* ctx, int kind, const int16_t* dOrg, int orgStride, int64_t orgBlockStride,
                     const int16_t* dCur, int curStride, int64_t curBlockStride, int w, int h, int subShift, int n,
                     uint64_t* dOut);
/* One block pair in HOST memory, synchronous: the signature a DistParam hook needs (RdCost.h:60-105). */
int vtmme_dist_host(vtmme_ctx* ctx, int kind, const int16_t* org, int orgStride, const int16_t* cur, int curStride,
                    int w, int h, int subShift, uint64_t* out);

/* Separable interpolation of n blocks (DEVICE pointers), semantics of InterpolationFilter::filterHor /
 * filterVer (InterpolationFilter.cpp:749-895): comp 0 luma 8-tap (frac in 1/16), comp 1 chroma 4-tap of
 * 4:2:0 (frac in 1/32); vertical 0/1; isFirst/isLast as in the reference; src points at the first
 * output sample position (the callee steps back (N/2-1) taps itself, InterpolationFilter.cpp:574-575). */
int vtmme_interp_batch(vtmme_ctx* ctx, int comp, int vertical, const int16_t* dSrc, int srcStride,
                       int64_t srcBlockStride, int16_t* dDst, int dstStride, int64_t dstBlockStride, int w, int h,
                       int frac, int isFirst, int isLast, int bitDepth, int useAltHpel, int n);
int vtmme_interp_host(vtmme_ctx* ctx, int comp, int vertical, const int16_t* src, int srcStride, int16_t* dst,
                      int dstStride, int w, int h, int frac, int isFirst, int isLast, int bitDepth, int useAltHpel);

/* The function-pointer flavour of the same filters: explicit taps, exactly the arguments the reference's table
 * entries m_filterHor[taps][isFirst][isLast] / m_filterVer / m_filterCopy receive (InterpolationFilter.h:96-98).
 * nTaps 8, 4 or 2; copy != 0 selects filterCopy<isFirst,isLast> (coeff ignored).  HOST pointers, synchronous. */
int vtmme_filter_host(vtmme_ctx* ctx, int nTaps, int vertical, int isFirst, int isLast, int copy, const int16_t* src,
                      int srcStride, int16_t* dst, int dstStride, int w, int h, const int16_t* coeff, int bitDepth);

/* ---- motion compensation (the consumer of the motion vectors the search returns) ------------------
 * Uni-directional block prediction with the semantics of InterPrediction::xPredInterBlk
 * (CommonLib/InterPrediction.cpp:660-830), plain path: no RPR, wrap-around, BDOF padding, DMVR or
 * bilinear filter.  comp 0: luma plane, 8-tap, 4 fractional MV bits; comp 1: a 4:2:0 chroma plane
 * (uploaded as a picture of its own), 4-tap, 5 fractional bits of the SAME luma MV (:675-676).
 * bi != 0 keeps the 14-bit intermediates (rndRes = !bi, :673) for vtmme_add_avg.  The MV must already
 * be clipped (clipMv, as xPredInterUni does): a block whose taps leave the padded plane is
 * VTMME_ERR_RANGE.  Block i's w*h samples are written packed (row stride w) at the sum of the
 * previous blocks' sizes. */
typedef struct vtmme_mc_block
{
  int32_t refPic;    /* uploaded plane of the component */
  int32_t x, y;      /* block position in samples of that plane */
  int32_t w, h;      /* 1..128 */
  int32_t mvX, mvY;  /* 1/16 luma sample (MV_PRECISION_INTERNAL) */
  int32_t reserved;  /* 0 */
} vtmme_mc_block;

/* blocks: HOST array; dDst: DEVICE buffer; asynchronous on the context stream. */
int vtmme_mc_batch(vtmme_ctx* ctx, int comp, int bi, int bitDepth, int useAltHpel, int n, const vtmme_mc_block* blocks,
                   int16_t* dDst);
/* dst: HOST buffer; synchronous. */
int vtmme_mc_host(vtmme_ctx* ctx, int comp, int bi, int bitDepth, int useAltHpel, int n, const vtmme_mc_block* blocks,
                  int16_t* dst);
/* AreaBuf<Pel>::addAvg (CommonLib/Buffer.cpp:467-507): dDst = clip((dSrc0 + dSrc1 + offset) >> shift) over count
 * samples of two bi != 0 predictions.  DEVICE pointers, asynchronous. */
int vtmme_add_avg(vtmme_ctx* ctx, const int16_t* dSrc0, const int16_t* dSrc1, int16_t* dDst, int64_t count, int bitDepth);
/* AreaBuf<T>::removeHighFreq (CommonLib/Buffer.h:474-517): dOrg = 2*dOrg - dPred (clipped to the sample range when
 * clip != 0) — how the pattern of a bi-predictive search is formed from the original block and the other
 * direction's prediction (InterSearch.cpp:3317-3325).  DEVICE pointers, asynchronous. */
int vtmme_remove_high_freq(vtmme_ctx* ctx, int16_t* dOrg, const int16_t* dPred, int64_t count, int clip, int bitDepth);
/* BCW (bi-prediction with CU-level weights, `BCW : 1` in the random-access and low-delay-B configurations):
 * AreaBuf<Pel>::addWeightedAvg (CommonLib/Buffer.cpp:87-121): dDst = clip((dSrc0*w0 + dSrc1*w1 + offset) >> shift) with
 * w1 = g_BcwWeights[bcwIdx] = {-2, 3, 4, 5, 10}, w0 = 8 - w1;
 * AreaBuf<Pel>::removeWeightHighFreq (CommonLib/Buffer.h:78-79,124; Buffer.cpp:123-140): dOrg = (dOrg*weight0 -
 * dPred*weight1 + 2^15) >> 16 with normalizer = ((1 << 16) + |bcwWeight| / 2) / bcwWeight, weight0 = 8 * normalizer,
 * weight1 = (8 - bcwWeight) * normalizer — the pattern of a bi-predictive search under a BCW weight
 * (InterSearch.cpp:3317-3325 with getBcwWeight).  DEVICE pointers, asynchronous. */
int vtmme_add_weighted_avg(vtmme_ctx* ctx, const int16_t* dSrc0, const int16_t* dSrc1, int16_t* dDst, int64_t count, int bitDepth,
                           int bcwIdx);
int vtmme_remove_weight_high_freq(vtmme_ctx* ctx, int16_t* dOrg, const int16_t* dPred, int64_t count, int clip, int bitDepth,
                                  int bcwWeight);

/* ---- candidate distortion (the step BEFORE the search: AMVP template cost and ME seeds) ----------
 * For one PU and a list of candidate MVs: SAD(org, xPredInterBlk(ref, mv)) per candidate, without
 * materialising the predictions.  With fractional MVs and subShift 0 this is the distortion term of
 * InterSearch::xGetTemplateCost (InterSearch.cpp:3235-3270, getDistPart DF_SAD); with integer MVs
 * (multiples of 16) and the search's subShift it is the SAD of the predictor / history-MV seeds that
 * choose the search-window centre (InterSearch.cpp:3388-3426).  The caller adds its own rate term
 * (m_auiMVPIdxCost resp. getCostOfVectorWithPredictor) and keeps the first strict minimum.  MVs must be
 * clipped already (clipMv).  out: HOST array, one uint64 per candidate in job order.  Synchronous. */
typedef struct vtmme_cand_job
{
  int32_t        curPic;     /* picture holding the original block; ignored when org != NULL           */
  int32_t        refPic;
  int32_t        x, y, w, h; /* PU luma rectangle, w,h in [1,128]                                      */
  const int16_t* org;        /* optional HOST pointer to the pattern (row stride orgStride)            */
  int32_t        orgStride;
  int32_t        nCand;      /* 1..64                                                                  */
  const int32_t* mv;         /* HOST: nCand x {mvX, mvY}, 1/16 luma sample                             */
  int32_t        subShift;   /* DistParam::subShift                                                    */
  int32_t        reserved;   /* 0 */
} vtmme_cand_job;
int vtmme_cand_sad(vtmme_ctx* ctx, int bitDepth, int useAltHpel, int nJobs, const vtmme_cand_job* jobs, uint64_t* out);

/* ---- decoder-side MV refinement (DMVR), SURVEY 8(f) rank 4 ---------------------------------------------------------
 * The search of InterPrediction::xProcessDMVR for a batch of sub-blocks (CommonLib/InterPrediction.cpp:2098-2154):
 * bilinear prediction of both lists around the merge MVs (xPrefetch :1664-1708, xinitMC :1949-1995), the cost at the
 * merge MVs with its 1/4 bias and the early exit below w*h, the 25 mirrored integer offsets (xDMVRCost :1919-1927,
 * xBIPMVRefine :1820-1843) and the parametric sub-sample step (xDMVRSubPixelErrorSurface :1929-1947).  The result is
 * what xProcessDMVR stores in pu.mvdL0SubPu[] (list 0 moves by +mvd, list 1 by -mvd); the final motion compensation
 * with the refined MVs is vtmme_mc_batch's job.  Plain DMVR only: no wrap-around, no reference scaling. */
typedef struct
{
  int32_t x, y, w, h;         /* sub-block (luma samples); w, h in {8, 16} (DMVR_SUBCU_WIDTH / HEIGHT, PU::checkDMVRCondition) */
  int32_t mvL0x, mvL0y;       /* merge MV of list 0, 1/16 sample                                                      */
  int32_t mvL1x, mvL1y;       /* merge MV of list 1                                                                   */
} vtmme_dmvr_block;

typedef struct
{
  int32_t  mvdX, mvdY;        /* pu.mvdL0SubPu[i], 1/16 sample                                                        */
  uint32_t minCost;           /* minCost after the integer stage (the BDOF switch compares it with 2*w*h, :2146)      */
  int32_t  notZeroCost;       /* 0: the search stopped at the centre (cost below w*h, or zero)                        */
} vtmme_dmvr_result;

/* refPic0 / refPic1: uploaded reference pictures of list 0 / list 1 (same size); blocks, results: HOST arrays of n entries;
 * maxCu: sps.getMaxCUWidth() (== Height) of the MV clip, <= 128; n <= 2^20 (an 8K picture has 129,600 sub-blocks).
 * Synchronous. */
int vtmme_dmvr_refine(vtmme_ctx* ctx, int refPic0, int refPic1, int bitDepth, int maxCu, int n, const vtmme_dmvr_block* blocks,
                      vtmme_dmvr_result* results);

/* ---- affine motion estimation: AffineGradientSearch's dispatch-table primitives ------------------------------
 * The three function pointers every iteration of InterSearch::xAffineMotionEstimation calls (EncoderLib/InterSearch.cpp:5486-5526;
 * CommonLib/AffineGradientSearch.h:52-54, scalar :64-174, SIMD x86/AffineGradientSearchX86.h):
 *   m_HorizontalSobelFilter / m_VerticalSobelFilter   3x3 Sobel of the prediction into int derivatives, border = nearest interior value
 *   m_EqualCoeffComputer                              sums of products of the per-sample model vector (4- or 6-parameter) and of its
 *                                                     product with the residual (<< 3), int64, accumulated into coeff[7][7]
 * Table flavour (HOST pointers, one block per call — what initAffineGradientSearchCUDA installs, cf. initRdCostCUDA): */
int vtmme_affine_sobel_host(vtmme_ctx* ctx, int vertical, const int16_t* pred, int predStride, int w, int h, int32_t* deriv, int derivStride);
int vtmme_affine_equal_coeff_host(vtmme_ctx* ctx, const int16_t* residue, int residueStride, const int32_t* d0, const int32_t* d1,
                                  int derivStride, int w, int h, int sixParam, int64_t* coeff /* [7][7], accumulated */);
/* Batched flavour: one gradient step — error = org - pred, both derivatives, the sums — for n blocks in ONE launch.
 * coeff: HOST, n * 49 entries, block i's matrix (rows 1..p, columns 0..p; p = 4 or 6) starting from zero. */
typedef struct
{
  const int16_t* org;        /* HOST: original block (for bi-prediction: 2*org - other prediction)  */
  int32_t        orgStride;
  const int16_t* pred;       /* HOST: current affine prediction of the block (xPredAffineBlk)       */
  int32_t        predStride;
  int32_t        w, h;       /* 4..128                                                              */
  int32_t        sixParam;   /* cu.affineType == AFFINEMODEL_6PARAM                                 */
  int32_t        reserved;   /* 0                                                                   */
} vtmme_affine_block;
int vtmme_affine_gradient_step(vtmme_ctx* ctx, int n, const vtmme_affine_block* blocks, int64_t* coeff);

/* ---- symmetric-MVD search -------------------------------------------------------------------------
 * InterSearch::xSymmetricMotionEstimation (EncoderLib/InterSearch.cpp:4506-4518): the MV of the searched list moves on a
 * diamond (at most 8 >> imv rounds) and once on a cross (xSymmeticRefineMvSearch :4393-4503), the MV of the other list
 * mirrors the MV difference; every candidate costs the MVD rate plus xGetSymmetricCost (:4341-4391): both 8-tap
 * predictions at the clipped MVs, 2*org - predCur (removeHighFreq, or removeWeightHighFreq under a BCW weight), SATD
 * (HadamardME) or SAD against predTar, weighted by xGetMEDistortionWeight.  MVs in 1/16 sample. */
typedef struct
{
  int32_t        curPic;               /* uploaded original picture (ignored when org != NULL)                        */
  int32_t        refPicCur, refPicTar; /* uploaded reference pictures of the searched list and of the other list       */
  int32_t        x, y, w, h;           /* PU luma rectangle, w and h powers of two in 8..128                           */
  const int16_t* org;                  /* optional HOST pattern (w x h at orgStride) instead of the block of curPic    */
  int32_t        orgStride;
  int32_t        maxCu;                /* sps.getMaxCUWidth() of the MV clip (clipMvInPic), <= 128                     */
  int32_t        bitDepth;
  int32_t        imv;                  /* cu.imv: 0 quarter, 1 integer, 2 four-sample, 3 half sample (alt. filter)     */
  int32_t        curPredX, curPredY;   /* rcMvCurPred                                                                  */
  int32_t        tarPredX, tarPredY;   /* rcMvTarPred                                                                  */
  int32_t        curMvX, curMvY;       /* rCurMvField.mv on entry                                                      */
  int32_t        tarMvX, tarMvY;       /* rTarMvField.mv on entry                                                      */
  int32_t        clipBiPred;           /* EncCfg::getClipForBiPredMeEnabled()                                          */
  int32_t        useHad;               /* !slice->getDisableSATDForRD()                                                */
  int32_t        bcwIdx;               /* cu.BcwIdx, 0..4 (2 = BCW_DEFAULT)                                            */
  double         lambdaMotion;         /* RdCost::m_motionLambda (selectMotionLambda)                                  */
  uint64_t       cost;                 /* ruiCost on entry                                                             */
} vtmme_smvd;

typedef struct
{
  int32_t  curMvX, curMvY, tarMvX, tarMvY;   /* rCurMvField.mv / rTarMvField.mv on return */
  uint64_t cost;                             /* ruiCost on return                         */
} vtmme_smvd_result;

/* jobs, results: HOST arrays of n entries (n <= 65536); one launch for all of them.  Synchronous. */
int vtmme_smvd_search(vtmme_ctx* ctx, int n, const vtmme_smvd* jobs, vtmme_smvd_result* results);

/* The prediction of ONE list after DMVR — InterPrediction::xFinalPaddedMCForDMVR (CommonLib/InterPrediction.cpp:1845-1917) over the
 * buffer xPrefetch (:1664-1708) filled and xPad (:1710-1730) extended: the 8-tap (comp 0) / 4-tap (comp 1, 4:2:0 chroma
 * plane uploaded as a picture of its own) filter of xPredInterBlk with bi = true (14-bit intermediates for
 * vtmme_add_avg) at the fraction of clip(refined MV), reading the (w + taps - 1) x (h + taps - 1) window at the integer
 * part of clip(merge MV - (taps/2 - 1) samples) with replicated borders instead of the picture.  blocks[i]: {x, y, w, h} =
 * the LUMA rectangle of the sub-block, mvL0 = the list's merge MV, mvL1 = the list's refined MV (merge MV +- mvdL0SubPu),
 * both 1/16 luma sample.  A chroma block that did not move is the plain vtmme_mc_batch prediction (the reference reads
 * the picture then, :1871).  dst: HOST, predictions packed (block i at the sum of earlier (w*h) >> (2*comp)).
 * Synchronous. */
int vtmme_dmvr_final_mc(vtmme_ctx* ctx, int comp, int refPic, int bitDepth, int maxCu, int n, const vtmme_dmvr_block* blocks,
                        int16_t* dst);

/* ---- GOP-based temporal filter: motion estimation ---------------------------------
 * EncTemporalFilter::motionEstimation (EncoderLib/EncTemporalFilter.cpp:448-466): the four-level hierarchical block
 * search of a reference frame against the original (16x16 blocks on the 1/4, 1/2 and full resolution pictures, then
 * 8x8 blocks with a 1/16-sample refinement; squared error, 6-tap interpolation — motionEstimationLuma :363-446,
 * motionErrorLuma :268-361, subsampleLuma :241-266).  Pair i searches picture refPics[i] ("buffer") against
 * orgPics[i]; both uploaded luma planes of the same size (the library replicates / keeps their border; the filter's
 * own padding is 128 samples).  mv: HOST array, per pair (height/4) x (width/4) entries {x, y, error} (x, y in 1/16
 * sample, row stride width/4 entries) exactly as TemporalFilterSourcePicInfo::mvs is filled (:205): block (bx, by) of
 * the 8x8 level at entry (bx, by), entries no block writes {0, 0, INT32_MAX}.  Synchronous. */
int vtmme_mctf_me(vtmme_ctx* ctx, int nPairs, const int32_t* orgPics, const int32_t* refPics, int bitDepth, int32_t* mv);
/* EncTemporalFilter::applyMotion (EncTemporalFilter.cpp:470-552) for one component: the uploaded plane srcPic
 * (luma, or a chroma plane uploaded as a picture of its own with csx / csy = 1 for 4:2:0) motion-compensated block
 * by block — (8 >> csx) x (8 >> csy) samples per luma 8x8 block — with the vectors of vtmme_mctf_me (HOST, mvStride
 * entries per row).  dst: HOST, plane width x height samples, packed; samples outside whole blocks are returned 0
 * (the reference leaves them unwritten).  Synchronous. */
int vtmme_mctf_apply_motion(vtmme_ctx* ctx, int srcPic, int csx, int csy, const int32_t* mv, int mvStride, int mvRows,
                            int bitDepth, int16_t* dst);

/* EncTemporalFilter::bilateralFilter (EncTemporalFilter.cpp:555-622), the weighting of one component: every sample of the
 * uploaded original plane orgPic is replaced by the weighted mean of itself (weight 1) and the co-located samples of the
 * numRefs (<= 8) uploaded planes corrPics[i] — the neighbouring pictures after vtmme_mctf_apply_motion.  The weight of a
 * neighbouring sample depends only on |refVal - orgVal| and the picture: weights = HOST array of numRefs tables of
 * (1 << bitDepth) doubles, weights[i][d] = weightScaling * m_refStrengths[row][min(1, |origOffset_i| - 1)] *
 * exp(-(d * 1024 / 2^bitDepth)^2 / (2 * sigmaSq)) as the caller's own libm evaluates it (:595-610) — the device then only
 * multiplies, adds and divides IEEE doubles in the reference's order, and every output sample equals the reference's.
 * dst: HOST, width x height samples, packed.  Synchronous. */
int vtmme_mctf_bilateral(vtmme_ctx* ctx, int orgPic, int numRefs, const int32_t* corrPics, const double* weights, int bitDepth,
                         int16_t* dst);

/* ---- measurement helpers ------------------------------------------------------------------------
 * Per-kernel timing of the frame path: when enabled, vtmme_search_frames[_device] brackets each of its
 * kernels with CUDA events on the context stream; vtmme_frame_kernel_ms returns the durations of the most
 * recent call, in launch order: [0] me_tree_sad (8/16/32 levels), [1] me_tree_upper (64/128), [2] me_frac_frame.
 * It synchronises the stream. */
int vtmme_set_profiling(vtmme_ctx* ctx, int enable);
int vtmme_frame_kernel_ms(vtmme_ctx* ctx, float ms[3]);

/* ------------------------------------------------------------------------------------------------
 * Issue-rate microbenchmark of one instruction class on the current device (roofline denominator of
 * the integer search).  variant: see vtm_b200/peaks.py. */
int vtmme_int_peak(int variant, int iters, double* laneInstrPerClkPerSm, double* ms, double* smClockMHz);

#ifdef __cplusplus
}
#endif
#endif
