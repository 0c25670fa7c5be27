/* TEST INFRASTRUCTURE — CPU oracle, not part of the product.
 *
 * Plain-C restatement of the arithmetic of VTM 9.3's inter motion-search hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it;
 * the product (libvtmme.so) never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors of its own, so the
 * oracle is pinned against outputs of the UNMODIFIED reference compiled here (oracle/_ref/libvtmref.so,
 * built by oracle/Makefile.ref from /root/reference) in tests/test_oracle_vs_ref.py, and against
 * committed fixtures generated from it (tests/golden/, generator tests/golden/make_golden.py).
 */
#ifndef VTM_ORACLE_H
#define VTM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int16_t vo_pel; /* Pel, TypeDef.h:259 */

/* Same field layout as RefSearchJob / RefSearchResult in oracle/ref_harness.cpp. */
typedef struct
{
  const vo_pel* org;
  int           orgStride;
  int           w, h;
  const vo_pel* refAtPU; /* cStruct.piRefY: reference plane at the PU position */
  int           refStride;
  int           srLeft, srRight, srTop, srBottom;
  int           predQx, predQy; /* quarter-pel AMVP predictor */
  int           imvShift;
  int           subShiftMode;
  int           bitDepth;
  int           useHad;
  int           useAltHpel;
  int           doFrac;
  double        lambdaMotion;
} vo_job;

typedef struct
{
  int      mvX, mvY;
  uint64_t intSad;
  int      halfX, halfY;
  int      qterX, qterY;
  uint64_t fracCost;
} vo_result;

/* distortion */
int      vo_subshift(int subShiftMode, int w, int h);
uint64_t vo_sad(const vo_pel* org, int orgStride, const vo_pel* cur, int curStride, int w, int h, int subShift);
uint64_t vo_satd(const vo_pel* org, int orgStride, const vo_pel* cur, int curStride, int w, int h);

/* motion-vector rate */
uint32_t vo_eg_bits(int v);
uint32_t vo_mv_bits(int x, int y, int predX, int predY, int costScale, int imvShift);
uint64_t vo_mv_cost(double lambdaMotion, uint32_t bits);

/* interpolation (comp: 0 luma, 1/2 chroma of 4:2:0); frac in 1/16 (luma) or 1/32 (chroma) units */
void vo_filter_hor(int comp, const vo_pel* src, int srcStride, vo_pel* dst, int dstStride, int w, int h, int frac,
                   int isLast, int bd, int useAltHpel);
void vo_filter_ver(int comp, const vo_pel* src, int srcStride, vo_pel* dst, int dstStride, int w, int h, int frac,
                   int isFirst, int isLast, int bd, int useAltHpel);

/* search window (xSetSearchRange) — pred in 1/16-pel internal units */
void vo_clip_mv(int* mvx, int* mvy, int posX, int posY, int picW, int picH, int maxCuW, int maxCuH);
void vo_set_search_range(int predX16, int predY16, int posX, int posY, int picW, int picH, int maxCuW, int maxCuH,
                         int searchRange, int* left, int* right, int* top, int* bottom);

/* integer full search (xPatternSearch) */
void vo_pattern_search(const vo_job* j, int* mvx, int* mvy, uint64_t* sad);

/* fractional refinement (xPatternSearchFracDIF): literal buffers and direct form */
void vo_frac_literal(const vo_job* j, int mvx, int mvy, int* hx, int* hy, int* qx, int* qy, uint64_t* cost);
void vo_frac_direct(const vo_job* j, int mvx, int mvy, int* hx, int* hy, int* qx, int* qy, uint64_t* cost);
/* one prediction block at quarter-pel offset (dqx,dqy) from integer MV (mvx,mvy), direct form */
void vo_pred_qpel(const vo_job* j, int mvx, int mvy, int dqx, int dqy, int useAltHpel, vo_pel* dst, int dstStride);

/* xPatternSearch + xPatternSearchFracDIF body; literal!=0 picks the literal fractional path */
void vo_search(const vo_job* j, vo_result* r, int literal);
double vo_search_batch(const vo_job* jobs, vo_result* res, int n, int literal);

/* motion compensation of one block (xPredInterBlk plain path), bi-pred average, bi-pred search target */
void vo_mc_block(int comp, const vo_pel* refAtBlk, int refStride, int w, int h, int mvX, int mvY, int bi, int bd,
                 int useAltHpel, vo_pel* dst, int dstStride);
void vo_add_avg(const vo_pel* s0, const vo_pel* s1, vo_pel* dst, int n, int bd);
void vo_add_weighted_avg(const vo_pel* s0, const vo_pel* s1, vo_pel* dst, int n, int bd, int bcwIdx);
void vo_remove_high_freq(vo_pel* dst, const vo_pel* src, int n, int clip, int bd);
void vo_remove_weight_high_freq(vo_pel* dst, const vo_pel* src, int n, int clip, int bd, int bcwWeight);

/* tail of xMotionEstimation (InterSearch.cpp:3477-3484): final MV (quarter-pel), bits and cost */
void vo_me_finish(const vo_job* j, const vo_result* r, double fWeight, uint32_t bitsIn, int* mvQx, int* mvQy,
                  uint32_t* bitsOut, uint64_t* costOut);

/* Integer-pel / 4-pel AMVR refinement, InterSearch::xPatternSearchIntRefine (InterSearch.cpp:4172-4282).
 * Same field layout as RefIntRefine in oracle/ref_harness.cpp.  MVs in 1/16 sample (MV_PRECISION_INTERNAL). */
typedef struct
{
  int      imv;            /* cu.imv: 1 IMV_FPEL, 2 IMV_4PEL                                            */
  int      mvX, mvY;       /* in: rcMv (integer-pel precise); out: refined MV                            */
  int      numCand;        /* amvpInfo.numCand (1 or 2)                                                  */
  int      candX[2], candY[2]; /* amvpInfo.mvCand[] (already rounded to the AMVR precision)              */
  int      mvpIdx;         /* in/out riMVPIdx                                                            */
  uint32_t mvpIdxBits[2];  /* m_auiMVPIdxCost[i][AMVP_MAX_NUM_CANDS]                                     */
  uint32_t bits;           /* in/out ruiBits                                                             */
  double   fWeight;        /* xGetMEDistortionWeight                                                     */
  int      posX, posY, picW, picH, maxCuW, maxCuH; /* clipMv (clipMvInPic) arguments                    */
  uint64_t cost;           /* out ruiCost                                                                */
} vo_int_refine_io;
void vo_int_refine(const vo_job* j, vo_int_refine_io* io);

/* TZ search, InterSearch::xTZSearch (InterSearch.cpp:3640-3974) with its helpers xTZSearchHelp (:330-417),
 * xTZ2PointSearch (:420-446) and xTZ8PointDiamondSearch (:503-705): the integer search of FastSearch=1
 * (MESEARCH_DIAMOND: extended 0, fast 0), FastSearch=3 (MESEARCH_DIAMOND_ENHANCED: extended 1) and of the cached-MV
 * re-search (fast 1, :3445).  Same field layout as RefTzParams in oracle/ref_harness.cpp.  No hash ME, no MCTS, no
 * composite reference; subShiftMode 0, 1 or 2. */
typedef struct
{
  int startX, startY;       /* rcMv on entry, 1/16 sample                                                  */
  int hasInt2Nx2N;          /* pIntegerMv2Nx2NPred != NULL                                                 */
  int int2Nx2NX, int2Nx2NY; /*   its value, integer pel                                                    */
  int nSeeds;               /* m_uniMvListSize (<= 15)                                                     */
  int seedX[16], seedY[16]; /* uniMvs[list][ref] of the history entries, newest first, 1/16 sample         */
  int searchRange;          /* m_iSearchRange                                                              */
  int extended, fast;       /* bExtendedSettings, bFastSettings                                            */
  int firstSearchStop;      /* EncCfg::getFastMEAssumingSmootherMVEnabled                                  */
  int posX, posY, picW, picH, maxCuW, maxCuH;
  int selective;            /* 1: xTZSearchSelective (:3979-4170, FastSearch=2) instead of xTZSearch; extended, fast and  */
                            /*    firstSearchStop are then unused.  With vo_job.subShiftMode 1 the probes of either search */
                            /*    use xTZSearchHelp's staged SAD (:340-391)                                                 */
} vo_tz_params;
/* mvx,mvy: rcMv (integer pel); sad: ruiSAD; nProbes (optional): number of xTZSearchHelp + seed distortions */
void vo_tz_search(const vo_job* j, const vo_tz_params* p, int* mvx, int* mvy, uint64_t* sad, int* nProbes);

/* GOP-based temporal filter: hierarchical motion estimation of one reference frame against the original,
 * EncTemporalFilter::motionEstimation (EncoderLib/EncTemporalFilter.cpp:448-466) with motionEstimationLuma (:363-446),
 * motionErrorLuma (:268-361) and subsampleLuma (:241-266).
 * org / ref: sample (0,0) of planes that carry at least 128 replicated border samples on every side.
 * mv: (width/4) x (height/4) entries of {x, y, error} (x, y in 1/16 sample), row stride width/4 entries, as
 * TemporalFilterSourcePicInfo::mvs is allocated (:205); block (bx, by) of the final 8x8 level is entry (bx, by);
 * entries no block writes keep {0, 0, INT32_MAX}. */
void vo_mctf_me(const vo_pel* org, int orgStride, const vo_pel* ref, int refStride, int width, int height, int bitDepth,
                int32_t* mv);
/* EncTemporalFilter::applyMotion (EncoderLib/EncTemporalFilter.cpp:470-552) for one component: the reference frame
 * motion-compensated block by block (8x8 luma blocks, i.e. (8>>csx) x (8>>csy) samples of the component) with the
 * vectors of vo_mctf_me.  src: sample (0,0) of the component's plane (>= 128>>cs border samples); mv: the luma field,
 * mvStride entries per row; dst: compW x compH samples, row stride dstStride (samples outside whole blocks untouched). */
void vo_mctf_apply_motion(const vo_pel* src, int srcStride, int compW, int compH, int csx, int csy, const int32_t* mv,
                          int mvStride, int bitDepth, vo_pel* dst, int dstStride);
/* one distortion probe (motionErrorLuma): block bs x bs at (x, y), displacement (dx, dy) in 1/16 sample */
int  vo_mctf_error(const vo_pel* org, int orgStride, const vo_pel* ref, int refStride, int x, int y, int dx, int dy, int bs,
                   int bestError, int bitDepth);

/* EncTemporalFilter::bilateralFilter (EncoderLib/EncTemporalFilter.cpp:555-623), the weighting of one component over the
 * motion-compensated neighbours (corrected[i], from vo_mctf_apply_motion); origOffset[i]: POC distance of neighbour i */
void vo_mctf_bilateral(const vo_pel* org, int orgStride, const vo_pel* const* corrected, int corrStride, const int* origOffset,
                       int numRefs, int w, int h, int isChroma, int qp, double overallStrength, int bitDepth, vo_pel* dst, int dstStride);

/* the same weights as a table over |refVal - orgVal| for one POC-distance class (index = min(1, |origOffset| - 1)) */
void vo_mctf_bilateral_weights(int isChroma, int qp, double overallStrength, int bitDepth, int numRefs, int index, double* table);

/* Symmetric MVD search, InterSearch::xSymmetricMotionEstimation (EncoderLib/InterSearch.cpp:4506-4518) with
 * xSymmeticRefineMvSearch (:4393-4503) and xGetSymmetricCost (:4341-4391); the searched list is list 0; no MCTS constraint.
 * Same field layout as RefSmvdIo in oracle/ref_harness.cpp. */
typedef struct
{
  int      x, y, w, h;                 /* PU (= CU) luma rectangle                                              */
  int      picW, picH, maxCuW, maxCuH; /* clipMv (clipMvInPic) arguments                                        */
  int      bd;                         /* internal bit depth                                                    */
  int      imv;                        /* cu.imv: 0 quarter, 1 integer, 2 four-sample, 3 half (alt. filter)     */
  int      curPredX, curPredY;         /* rcMvCurPred, 1/16 sample                                              */
  int      tarPredX, tarPredY;         /* rcMvTarPred                                                           */
  int      curMvX, curMvY;             /* in/out rCurMvField.mv                                                 */
  int      tarMvX, tarMvY;             /* in/out rTarMvField.mv                                                 */
  int      clipBiPred;                 /* EncCfg::getClipForBiPredMeEnabled                                     */
  int      useHad;                     /* !slice->getDisableSATDForRD()                                         */
  int      bcwIdx;                     /* cu.BcwIdx, 0..4 (2 = BCW_DEFAULT, equal weights)                       */
  double   lambda;                     /* RdCost::m_motionLambda                                                */
  uint64_t cost;                       /* in/out ruiCost                                                        */
} vo_smvd_io;
/* refCur / refTar: sample (0,0) of the reference planes of eRefPicList and of the other list (border extended) */
void vo_smvd_search(const vo_pel* org, int orgStride, const vo_pel* refCur, const vo_pel* refTar, int refStride, vo_smvd_io* io);

/* Decoder-side MV refinement, one sub-block of InterPrediction::xProcessDMVR (CommonLib/InterPrediction.cpp:2098-2154):
 * bilinear predictions of both lists around the merge MVs, cost at the centre, 25 mirrored integer offsets, parametric
 * sub-sample step.  ref0 / ref1: sample (0,0) of the two reference planes (border extended); w, h in {8, 16}; MVs in 1/16
 * sample.  out: {mvdL0SubPu.hor, .ver, minCost, notZeroCost}. */
void vo_dmvr_block(const vo_pel* ref0, const vo_pel* ref1, int refStride, int x, int y, int w, int h, int mv0x, int mv0y, int mv1x,
                   int mv1y, int picW, int picH, int maxCuW, int maxCuH, int bd, int32_t* out);

/* luma prediction (14-bit intermediates, bi = true) of one list after DMVR: xPrefetch + xPad + xFinalPaddedMCForDMVR
 * (CommonLib/InterPrediction.cpp:1664-1730, 1845-1917); merge MV and refined MV (merge +- mvdL0SubPu) in 1/16 sample */
void vo_dmvr_final_luma(const vo_pel* ref, int refStride, int x, int y, int w, int h, int mergeX, int mergeY, int refinedX,
                        int refinedY, int picW, int picH, int maxCuW, int maxCuH, int bd, vo_pel* dst);
/* the same for a 4:2:0 chroma plane of a MOVED block (x, y, w, h: the luma rectangle; ref: the chroma plane; dst stride w/2) */
void vo_dmvr_final_chroma(const vo_pel* ref, int refStride, int x, int y, int w, int h, int mergeX, int mergeY, int refinedX,
                          int refinedY, int picW, int picH, int maxCuW, int maxCuH, int bd, vo_pel* dst);

#ifdef __cplusplus
}
#endif
/* Affine ME: AffineGradientSearch's dispatch-table primitives (CommonLib/AffineGradientSearch.cpp:64-174) */
void vo_affine_sobel(int vertical, const vo_pel* pred, int predStride, int* deriv, int derivStride, int w, int h);
void vo_affine_equal_coeff(const vo_pel* residue, int residueStride, const int* d0, const int* d1, int derivStride, int64_t coeff[7][7],
                           int w, int h, int sixParam);

#endif
