// TEST INFRASTRUCTURE — not part of the product.
//
// Thin C-ABI shim over the UNMODIFIED reference (VTM 9.3) built by oracle/Makefile.ref from the
// sources under /root/reference.  It lets tests/ and bench.py's cpu_baseline / --impl reference leg
// call the reference's own functions for the motion-search hot path:
//
//   RdCost::setDistParam + DistParam::distFunc          (CommonLib/RdCost.cpp:238-324, x86/RdCostX86.h)
//   InterpolationFilter::filterHor / filterVer          (CommonLib/InterpolationFilter.cpp:749-895)
//   InterSearch::xPatternSearch                         (EncoderLib/InterSearch.cpp:3566-3608)
//   InterSearch::xExtDIFUpSamplingH/Q, xPatternRefinement (InterSearch.cpp:5840-6050, 707-761)
//   InterPrediction::xPredInterBlk                      (CommonLib/InterPrediction.cpp:660-830), AreaBuf::removeHighFreq
//
// xPatternSearchFracDIF itself needs a PredictionUnit with a slice; its PU-independent body
// (InterSearch.cpp:4296-4338) is driven here by calling the three reference functions it calls.
// Compiled with -fno-access-control so the protected members can be reached without touching
// the reference sources.  Output: oracle/_ref/libvtmref.so
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <chrono>
#include <atomic>

#include "CommonLib/CommonDef.h"
#include "CommonLib/AffineGradientSearch.h"
#include "CommonLib/RdCost.h"
#include "CommonLib/InterpolationFilter.h"
#include "CommonLib/InterPrediction.h"
#include "CommonLib/Buffer.h"
#include "CommonLib/Picture.h"
#include <deque>
#include "CommonLib/Slice.h"
#include "CommonLib/CodingStructure.h"
#include "EncoderLib/EncCfg.h"
#include "EncoderLib/InterSearch.h"
#include "EncoderLib/EncTemporalFilter.h"

namespace {

struct Probe : public InterSearch
{
  RdCost rd;
  EncCfg cfg;
  Probe()
  {
    InterPrediction::init(&rd, CHROMA_420, 128);
    m_pcEncCfg = &cfg;
    m_skipFracME = false;
    m_useCompositeRef = false;
  }
};

ClpRng makeClp(int bd)
{
  ClpRng c;
  c.min = 0;
  c.max = (1 << bd) - 1;
  c.bd = bd;
  c.n = 0;
  return c;
}

thread_local Probe* t_probe = nullptr;
Probe& probe()
{
  if (!t_probe) t_probe = new Probe();   // leaked on purpose: InterSearch::destroy() wants a full init
  return *t_probe;
}

}   // namespace

extern "C" {

// 0 SCALAR, 1 SSE41, 2 SSE42, 3 AVX, 4 AVX2, 5 AVX512 (CommonDef.h:609-617)
int ref_simd_level() { return (int) read_x86_extension_flags(); }

// ME-flavour setDistParam (RdCost.cpp:238-324) + call through the dispatch table.
uint64_t ref_dist(const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w, int h,
                  int bitDepth, int subShiftMode, int useHad)
{
  Probe& p = probe();
  DistParam dp;
  CPelBuf o(org, orgStride, w, h);
  p.rd.setDistParam(dp, o, cur, curStride, bitDepth, COMPONENT_Y, subShiftMode, 1, useHad != 0);
  return dp.distFunc(dp);
}

// subShift the reference picks for (mode,w,h)
int ref_subshift(int subShiftMode, int w, int h)
{
  Probe& p = probe();
  DistParam dp;
  static int16_t dummy[4];
  CPelBuf o(dummy, w, w, h);
  p.rd.setDistParam(dp, o, dummy, w, 10, COMPONENT_Y, subShiftMode, 1, false);
  return dp.subShift;
}

uint32_t ref_mv_bits(int x, int y, int predX, int predY, int costScale, int imvShift)
{
  Probe& p = probe();
  p.rd.setPredictor(Mv(predX, predY));
  p.rd.setCostScale(costScale);
  return p.rd.getBitsOfVectorWithPredictor(x, y, imvShift);
}

uint64_t ref_mv_cost(double lambdaMotion, uint32_t bits)
{
  Probe& p = probe();
  p.rd.m_motionLambda = lambdaMotion;
  return p.rd.getCost(bits);
}

void ref_filter_hor(int comp, const int16_t* src, int srcStride, int16_t* dst, int dstStride, int w, int h,
                    int frac, int isLast, int bd, int useAltHpel)
{
  Probe& p = probe();
  p.m_if.filterHor(ComponentID(comp), src, srcStride, dst, dstStride, w, h, frac, isLast != 0, CHROMA_420,
                   makeClp(bd), 0, false, useAltHpel != 0);
}

void ref_filter_ver(int comp, const int16_t* src, int srcStride, int16_t* dst, int dstStride, int w, int h,
                    int frac, int isFirst, int isLast, int bd, int useAltHpel)
{
  Probe& p = probe();
  p.m_if.filterVer(ComponentID(comp), src, srcStride, dst, dstStride, w, h, frac, isFirst != 0, isLast != 0,
                   CHROMA_420, makeClp(bd), 0, false, useAltHpel != 0);
}

struct RefSearchJob
{
  const int16_t* org;      // original block (uni-pred: in the original picture)
  int            orgStride;
  int            w, h;
  const int16_t* refAtPU;  // reference picture plane at the PU position (piRefY)
  int            refStride;
  int            srLeft, srRight, srTop, srBottom;   // integer-pel window (cStruct.searchRange)
  int            predQx, predQy;                     // AMVP predictor in quarter-pel (RdCost::setPredictor)
  int            imvShift;                           // 0 qpel, 1 hpel, 2 fpel, 4 4pel
  int            subShiftMode;                       // 0 or 2 (FEN)
  int            bitDepth;
  int            useHad;                             // HadamardME && !DisableSATDForRD
  int            useAltHpel;
  int            doFrac;                             // 0: integer only, 1: run the xPatternSearchFracDIF body
  double         lambdaMotion;
};

struct RefSearchResult
{
  int      mvX, mvY;         // best integer MV
  uint64_t intSad;           // ruiSAD of xPatternSearch (best − mv cost)
  int      halfX, halfY;     // rcMvHalf
  int      qterX, qterY;     // rcMvQter
  uint64_t fracCost;         // ruiCost after xPatternSearchFracDIF (SATD + mv cost)
};

// xPatternSearch (+ the PU-independent body of xPatternSearchFracDIF, InterSearch.cpp:4296-4338)
void ref_search(const RefSearchJob* j, RefSearchResult* r)
{
  Probe& p = probe();
  p.rd.m_motionLambda = j->lambdaMotion;
  p.rd.setPredictor(Mv(j->predQx, j->predQy));
  p.rd.setCostScale(2);
  p.cfg.setUseHADME(j->useHad != 0);
  p.m_lumaClpRng = makeClp(j->bitDepth);

  CPelBuf pattern(j->org, j->orgStride, j->w, j->h);
  InterSearch::IntTZSearchStruct cs;
  memset(&cs, 0, sizeof(cs));
  cs.pcPatternKey = &pattern;
  cs.piRefY       = j->refAtPU;
  cs.iRefStride   = j->refStride;
  cs.imvShift     = j->imvShift;
  cs.useAltHpelIf = j->useAltHpel != 0;
  cs.subShiftMode = j->subShiftMode;
  cs.searchRange.left   = j->srLeft;
  cs.searchRange.right  = j->srRight;
  cs.searchRange.top    = j->srTop;
  cs.searchRange.bottom = j->srBottom;

  Mv         mv;
  Distortion cost = 0;
  p.xPatternSearch(cs, mv, cost);
  r->mvX    = mv.hor;
  r->mvY    = mv.ver;
  r->intSad = cost;
  r->halfX = r->halfY = r->qterX = r->qterY = 0;
  r->fracCost = cost;
  if (!j->doFrac) return;

  // ---- body of xPatternSearchFracDIF (InterSearch.cpp:4296-4338), calling the reference's own helpers
  Mv  mvHalf, mvQter;
  int iOffset = mv.getHor() + mv.getVer() * cs.iRefStride;
  CPelBuf roi(cs.piRefY + iOffset, cs.iRefStride, *cs.pcPatternKey);
  if (cs.imvShift > IMV_FPEL)
  {
    p.rd.setDistParam(p.m_cDistParam, *cs.pcPatternKey, cs.piRefY + iOffset, cs.iRefStride, p.m_lumaClpRng.bd,
                      COMPONENT_Y, 0, 1, j->useHad != 0);
    cost = p.m_cDistParam.distFunc(p.m_cDistParam);
    cost += p.rd.getCostOfVectorWithPredictor(mv.getHor(), mv.getVer(), cs.imvShift);
    r->fracCost = cost;
    return;
  }
  p.rd.setCostScale(1);
  p.xExtDIFUpSamplingH(&roi, cs.useAltHpelIf);
  mvHalf = mv;
  mvHalf <<= 1;
  Mv baseRefMv(0, 0);
  cost = p.xPatternRefinement(cs.pcPatternKey, baseRefMv, 2, mvHalf, true);
  if (cs.imvShift == IMV_OFF)
  {
    p.rd.setCostScale(0);
    p.xExtDIFUpSamplingQ(&roi, mvHalf);
    baseRefMv = mvHalf;
    baseRefMv <<= 1;
    mvQter = mv;
    mvQter <<= 1;
    mvQter += mvHalf;
    mvQter <<= 1;
    cost = p.xPatternRefinement(cs.pcPatternKey, baseRefMv, 1, mvQter, true);
  }
  else
  {
    mvQter.setZero();
  }
  r->halfX    = mvHalf.hor;
  r->halfY    = mvHalf.ver;
  r->qterX    = mvQter.hor;
  r->qterY    = mvQter.ver;
  r->fracCost = cost;
}

// n block pairs stored back to back (block stride w*h), one distFunc call each; returns seconds (table-level baseline)
double ref_dist_batch(const int16_t* org, const int16_t* cur, int w, int h, int n, int bitDepth, int subShiftMode,
                      int useHad, uint64_t* out)
{
  Probe& p = probe();
  DistParam dp;
  CPelBuf o0(org, w, w, h);
  p.rd.setDistParam(dp, o0, cur, w, bitDepth, COMPONENT_Y, subShiftMode, 1, useHad != 0);
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; i++)
  {
    dp.org.buf = org + (size_t) i * w * h;
    dp.cur.buf = cur + (size_t) i * w * h;
    out[i]     = dp.distFunc(dp);
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// n blocks: source blocks (w+8)x(h+8) back to back (output position at +4,+4), destination blocks w x h back to back
double ref_filter_batch(int comp, int vertical, const int16_t* src, int16_t* dst, int w, int h, int n, int frac,
                        int isFirst, int isLast, int bd)
{
  Probe&       p  = probe();
  const ClpRng c  = makeClp(bd);
  const int    sw = w + 8, sh = h + 8;
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; i++)
  {
    const int16_t* s = src + (size_t) i * sw * sh + 4 * sw + 4;
    int16_t*       d = dst + (size_t) i * w * h;
    if (vertical)
      p.m_if.filterVer(ComponentID(comp), s, sw, d, w, w, h, frac, isFirst != 0, isLast != 0, CHROMA_420, c);
    else
      p.m_if.filterHor(ComponentID(comp), s, sw, d, w, w, h, frac, isLast != 0, CHROMA_420, c);
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---- motion compensation: the reference's own InterPrediction::xPredInterBlk (InterPrediction.cpp:660-830) -------------
// A Picture is created with the given luma size, the reconstruction plane of component `comp` (0 Y, 1 Cb) is loaded from
// `plane` (component-sized, `margin` border samples on each side already extended), and xPredInterBlk runs for every block.
// blk: n x 6 int32 {x, y, w, h, mvX, mvY}; x,y,w,h in samples of the component, mv in 1/16 luma sample.
// dst: blocks packed back to back (row stride = w); *seconds (optional) = time spent in the xPredInterBlk calls.  Returns 0, or -1 when the margin does not fit the Picture's own.
int ref_mc_blocks(int comp, const int16_t* plane, int planeStride, int lumaW, int lumaH, int margin, int n,
                  const int32_t* blk, int bi, int bitDepth, int imvHpel, int16_t* dst, double* seconds)
{
  Probe& p = probe();
  Picture pic;
  pic.create(CHROMA_420, Size(lumaW, lumaH), 128, 128 + 16, false, 0);
  pic.unscaledPic = &pic;
  PelBuf reco = pic.getRecoBuf(ComponentID(comp));
  const int cw = comp ? lumaW / 2 : lumaW, ch = comp ? lumaH / 2 : lumaH;
  if (margin > (int) (pic.margin >> (comp ? 1 : 0))) return -1;
  for (int y = -margin; y < ch + margin; y++)
    memcpy(reco.buf + (ptrdiff_t) y * reco.stride - margin, plane + (ptrdiff_t) (y + margin) * planeStride,
           sizeof(int16_t) * (cw + 2 * margin));

  PPS pps;
  pps.setPicWidthInLumaSamples(lumaW);
  pps.setPicHeightInLumaSamples(lumaH);
  SPS sps;
  // xPredInterBlk reads cs->sps / cs->pps only (wrap-around, RPR): a zeroed CodingStructure shell carries the two pointers
  std::vector<uint64_t> shell((sizeof(CodingStructure) + 7) / 8, 0);
  CodingStructure* cs = reinterpret_cast<CodingStructure*>(shell.data());
  cs->sps = &sps;
  cs->pps = &pps;
  const ClpRng clp = makeClp(bitDepth);
  const int    sc  = comp ? 1 : 0;
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; i++)
  {
    const int32_t* b = blk + 6 * i;
    const int w = b[2], h = b[3];
    CodingUnit     cu;
    cu.imv    = imvHpel ? IMV_HPEL : IMV_OFF;
    cu.affine = false;
    PredictionUnit pu(CHROMA_420, Area(b[0] << sc, b[1] << sc, w << sc, h << sc));
    pu.cu = &cu;
    pu.cs = cs;
    PelBuf     d(dst, w, w, h);
    PelBuf     none;
    PelUnitBuf out = comp ? PelUnitBuf(CHROMA_420, none, d, none) : PelUnitBuf(CHROMA_420, d);
    p.xPredInterBlk(ComponentID(comp), pu, &pic, Mv(b[4], b[5]), out, bi != 0, clp, false, false);
    dst += w * h;
  }
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  pic.destroy();
  return 0;
}

// ---- decoder-side MV refinement: the reference's own members InterPrediction::xPrefetch (InterPrediction.cpp:1664-1708),
// xinitMC (:1949-1995: bilinear prediction of the (w+4) x (h+4) neighbourhood of both lists), xDMVRCost (:1919-1927),
// xBIPMVRefine (:1820-1843) and the free function xDMVRSubPixelErrorSurface (:1929-1947), driven per sub-block the way
// xProcessDMVR does (:2098-2154).  plane0 / plane1: luma planes of the two reference pictures (`margin` border samples on
// each side already extended).  blk: n x 8 int32 {x, y, w, h, mvL0x, mvL0y, mvL1x, mvL1y}, w,h <= 16, MVs in 1/16 sample.
// out: n x 4 int32 {mvdL0SubPu.hor, .ver (1/16 sample), minCost, notZeroCost}.  Returns 0, or -1 when the margin does not fit.
extern "C++" void xDMVRSubPixelErrorSurface(bool notZeroCost, int16_t* totalDeltaMV, int16_t* deltaMV, uint64_t* pSADsArray);   // InterPrediction.cpp:1929
int ref_dmvr_blocks(const int16_t* plane0, const int16_t* plane1, int planeStride, int lumaW, int lumaH, int margin, int n,
                               const int32_t* blk, int bitDepth, int32_t* out)
{
  Probe&  p = probe();
  Picture pics[2];
  for (int l = 0; l < 2; l++)
  {
    Picture& pic = pics[l];
    pic.create(CHROMA_420, Size(lumaW, lumaH), 128, 128 + 16, false, 0);
    pic.unscaledPic = &pic;
    PelBuf reco = pic.getRecoBuf(COMPONENT_Y);
    if (margin > (int) pic.margin) return -1;
    const int16_t* plane = l ? plane1 : plane0;
    for (int y = -margin; y < lumaH + margin; y++)
      memcpy(reco.buf + (ptrdiff_t) y * reco.stride - margin, plane + (ptrdiff_t) (y + margin) * planeStride,
             sizeof(int16_t) * (lumaW + 2 * margin));
  }
  PPS pps;
  pps.setPicWidthInLumaSamples(lumaW);
  pps.setPicHeightInLumaSamples(lumaH);
  SPS sps;
  std::vector<uint64_t> shell((sizeof(CodingStructure) + 7) / 8, 0);
  CodingStructure* cs = reinterpret_cast<CodingStructure*>(shell.data());
  cs->sps = &sps;
  cs->pps = &pps;
  Slice* slice = new Slice();
  for (int l = 0; l < 2; l++)
  {
    slice->m_apcRefPicList[l][0] = &pics[l];
    slice->m_scalingRatio[l][0]  = SCALE_1X;
  }
  clipMv = clipMvInPic;
  ClpRngs clps;
  for (int c = 0; c < MAX_NUM_COMPONENT; c++) clps.comp[c] = makeClp(bitDepth);
  const int side = 2 * DMVR_NUM_ITERATION + 1;
  for (int i = 0; i < n; i++)
  {
    const int32_t* b = blk + 8 * i;
    const int dx = b[2], dy = b[3];
    CodingUnit cu;
    cu.imv    = IMV_OFF;
    cu.affine = false;
    cu.slice  = slice;
    PredictionUnit pu(CHROMA_420, Area(b[0], b[1], dx, dy));
    pu.cu = &cu;
    pu.cs = cs;
    pu.refIdx[0] = pu.refIdx[1] = 0;
    pu.mv[0] = Mv(b[4], b[5]);
    pu.mv[1] = Mv(b[6], b[7]);
    // buffers as xProcessDMVR sets them up (:2063-2078)
    p.m_biLinearBufStride = dx + 2 * DMVR_NUM_ITERATION;
    p.m_cYuvRefBuffDMVRL0 = PelUnitBuf(CHROMA_400, PelBuf(p.m_cRefSamplesDMVRL0[0], dx, dx, dy));
    p.m_cYuvRefBuffDMVRL1 = PelUnitBuf(CHROMA_400, PelBuf(p.m_cRefSamplesDMVRL1[0], dx, dx, dy));
    Pel* predL0 = p.m_cYuvPredTempDMVRL0 + DMVR_NUM_ITERATION * p.m_biLinearBufStride + DMVR_NUM_ITERATION;
    Pel* predL1 = p.m_cYuvPredTempDMVRL1 + DMVR_NUM_ITERATION * p.m_biLinearBufStride + DMVR_NUM_ITERATION;
    p.xPrefetch(pu, p.m_cYuvRefBuffDMVRL0, REF_PIC_LIST_0, 1);
    p.xPrefetch(pu, p.m_cYuvRefBuffDMVRL1, REF_PIC_LIST_1, 1);
    p.xinitMC(pu, clps);
    uint64_t sads[side * side];
    for (int k = 0; k < side * side; k++) sads[k] = MAX_UINT64;
    uint64_t* centre = sads + (side * side) / 2;
    int16_t   total[2] = { 0, 0 }, delta[2] = { 0, 0 };
    bool      notZero  = true;
    uint64_t  minCost  = p.xDMVRCost(bitDepth, predL0, p.m_biLinearBufStride, predL1, p.m_biLinearBufStride, dx, dy);
    minCost -= minCost >> 2;
    if (minCost < (uint64_t) (dx * dy))
      notZero = false;
    else
    {
      centre[0] = minCost;
      if (!minCost)
        notZero = false;
      else
      {
        p.xBIPMVRefine(bitDepth, predL0, predL1, minCost, delta, centre, dx, dy);
        total[0] = delta[0];
        total[1] = delta[1];
        centre += delta[1] * side + delta[0];
      }
    }
    total[0] = total[0] << MV_FRACTIONAL_BITS_INTERNAL;
    total[1] = total[1] << MV_FRACTIONAL_BITS_INTERNAL;
    xDMVRSubPixelErrorSurface(notZero, total, delta, centre);
    out[4 * i + 0] = total[0];
    out[4 * i + 1] = total[1];
    out[4 * i + 2] = (int32_t) minCost;
    out[4 * i + 3] = notZero;
  }
  delete slice;
  pics[0].destroy();
  pics[1].destroy();
  return 0;
}

// ---- luma prediction after DMVR: the reference's own xPrefetch, xPad and xFinalPaddedMCForDMVR (InterPrediction.cpp:1664-1730,
// 1845-1917) per sub-block, as xProcessDMVR drives them (:2156-2181).  blk as ref_dmvr_blocks; mvd: n x 2 (mvdL0SubPu, 1/16
// sample).  dst0 / dst1: the two lists' predictions (14-bit intermediates), blocks packed back to back (row stride = w).
// chroma0 / chroma1 (optional): Cb planes of the two pictures with margin/2 border samples, row stride chromaStride;
// dstC0 / dstC1: their predictions ((w/2) x (h/2) per block).
int ref_dmvr_final(const int16_t* plane0, const int16_t* plane1, int planeStride, const int16_t* chroma0, const int16_t* chroma1,
                   int chromaStride, int lumaW, int lumaH, int margin, int n, const int32_t* blk, const int32_t* mvd, int bitDepth,
                   int16_t* dst0, int16_t* dst1, int16_t* dstC0, int16_t* dstC1);
int ref_dmvr_final_luma(const int16_t* plane0, const int16_t* plane1, int planeStride, int lumaW, int lumaH, int margin, int n,
                        const int32_t* blk, const int32_t* mvd, int bitDepth, int16_t* dst0, int16_t* dst1)
{
  return ref_dmvr_final(plane0, plane1, planeStride, nullptr, nullptr, 0, lumaW, lumaH, margin, n, blk, mvd, bitDepth, dst0, dst1, nullptr, nullptr);
}
int ref_dmvr_final(const int16_t* plane0, const int16_t* plane1, int planeStride, const int16_t* chroma0, const int16_t* chroma1,
                   int chromaStride, int lumaW, int lumaH, int margin, int n, const int32_t* blk, const int32_t* mvd, int bitDepth,
                   int16_t* dst0, int16_t* dst1, int16_t* dstC0, int16_t* dstC1)
{
  Probe&  p = probe();
  Picture pics[2];
  for (int l = 0; l < 2; l++)
  {
    Picture& pic = pics[l];
    pic.create(CHROMA_420, Size(lumaW, lumaH), 128, 128 + 16, false, 0);
    pic.unscaledPic = &pic;
    PelBuf reco = pic.getRecoBuf(COMPONENT_Y);
    if (margin > (int) pic.margin) return -1;
    const int16_t* plane = l ? plane1 : plane0;
    for (int y = -margin; y < lumaH + margin; y++)
      memcpy(reco.buf + (ptrdiff_t) y * reco.stride - margin, plane + (ptrdiff_t) (y + margin) * planeStride,
             sizeof(int16_t) * (lumaW + 2 * margin));
    for (int c = 1; c < 3; c++)   // chroma is predicted too: defined samples (zero without chroma planes)
    {
      PelBuf cb = pic.getRecoBuf(ComponentID(c));
      for (int y = -(int) (pic.margin >> 1); y < lumaH / 2 + (int) (pic.margin >> 1); y++)
        memset(cb.buf + (ptrdiff_t) y * cb.stride - (pic.margin >> 1), 0, sizeof(int16_t) * (lumaW / 2 + pic.margin));
      const int16_t* cp = l ? chroma1 : chroma0;
      const int      cm = margin / 2;
      if (cp)
        for (int y = -cm; y < lumaH / 2 + cm; y++)
          memcpy(cb.buf + (ptrdiff_t) y * cb.stride - cm, cp + (ptrdiff_t) (y + cm) * chromaStride, sizeof(int16_t) * (lumaW / 2 + 2 * cm));
    }
  }
  PPS pps;
  pps.setPicWidthInLumaSamples(lumaW);
  pps.setPicHeightInLumaSamples(lumaH);
  SPS sps;
  std::vector<uint64_t> shell((sizeof(CodingStructure) + 7) / 8, 0);
  CodingStructure* cs = reinterpret_cast<CodingStructure*>(shell.data());
  cs->sps = &sps;
  cs->pps = &pps;
  Slice* slice = new Slice();
  for (int l = 0; l < 2; l++)
  {
    slice->m_apcRefPicList[l][0] = &pics[l];
    slice->m_scalingRatio[l][0]  = SCALE_1X;
  }
  for (int c = 0; c < MAX_NUM_COMPONENT; c++) slice->m_clpRngs.comp[c] = makeClp(bitDepth);
  cs->slice = slice;
  clipMv = clipMvInPic;
  PelStorage pred[2];
  for (int l = 0; l < 2; l++) pred[l].create(UnitArea(CHROMA_420, Area(0, 0, 16, 16)));
  for (int i = 0; i < n; i++)
  {
    const int32_t* b = blk + 8 * i;
    const int dx = b[2], dy = b[3];
    CodingUnit cu;
    cu.imv    = IMV_OFF;
    cu.affine = false;
    cu.slice  = slice;
    PredictionUnit pu(CHROMA_420, Area(b[0], b[1], dx, dy));
    pu.cu = &cu;
    pu.cs = cs;
    pu.refIdx[0] = pu.refIdx[1] = 0;
    const Mv mergeMv[2] = { Mv(b[4], b[5]), Mv(b[6], b[7]) };
    pu.mv[0] = mergeMv[0];
    pu.mv[1] = mergeMv[1];
    // the prefetch buffers as xProcessDMVR sets them up (:2063-2078), all three components
    p.m_cYuvRefBuffDMVRL0 = PelUnitBuf(CHROMA_420, PelBuf(p.m_cRefSamplesDMVRL0[0], dx, dx, dy), PelBuf(p.m_cRefSamplesDMVRL0[1], dx / 2, dx / 2, dy / 2),
                                       PelBuf(p.m_cRefSamplesDMVRL0[2], dx / 2, dx / 2, dy / 2));
    p.m_cYuvRefBuffDMVRL1 = PelUnitBuf(CHROMA_420, PelBuf(p.m_cRefSamplesDMVRL1[0], dx, dx, dy), PelBuf(p.m_cRefSamplesDMVRL1[1], dx / 2, dx / 2, dy / 2),
                                       PelBuf(p.m_cRefSamplesDMVRL1[2], dx / 2, dx / 2, dy / 2));
    p.xPrefetch(pu, p.m_cYuvRefBuffDMVRL0, REF_PIC_LIST_0, 1);
    p.xPrefetch(pu, p.m_cYuvRefBuffDMVRL1, REF_PIC_LIST_1, 1);
    const Mv   d(mvd[2 * i], mvd[2 * i + 1]);
    const bool moved = d != Mv(0, 0);
    if (moved)
    {
      p.xPrefetch(pu, p.m_cYuvRefBuffDMVRL0, REF_PIC_LIST_0, 0);
      p.xPrefetch(pu, p.m_cYuvRefBuffDMVRL1, REF_PIC_LIST_1, 0);
      p.xPad(pu, p.m_cYuvRefBuffDMVRL0, REF_PIC_LIST_0);
      p.xPad(pu, p.m_cYuvRefBuffDMVRL1, REF_PIC_LIST_1);
    }
    pu.mv[0] = mergeMv[0] + d;
    pu.mv[1] = mergeMv[1] - d;
    pu.mv[0].clipToStorageBitDepth();
    pu.mv[1].clipToStorageBitDepth();
    const UnitArea rel(CHROMA_420, Area(0, 0, dx, dy));
    PelUnitBuf s0 = pred[0].getBuf(rel), s1 = pred[1].getBuf(rel);
    p.xFinalPaddedMCForDMVR(pu, s0, s1, p.m_cYuvRefBuffDMVRL0, p.m_cYuvRefBuffDMVRL1, false, mergeMv, moved);
    for (int y = 0; y < dy; y++)
    {
      memcpy(dst0 + (ptrdiff_t) y * dx, s0.Y().buf + (ptrdiff_t) y * s0.Y().stride, sizeof(int16_t) * dx);
      memcpy(dst1 + (ptrdiff_t) y * dx, s1.Y().buf + (ptrdiff_t) y * s1.Y().stride, sizeof(int16_t) * dx);
    }
    dst0 += dx * dy;
    dst1 += dx * dy;
    if (dstC0 && dstC1)
    {
      for (int y = 0; y < dy / 2; y++)
      {
        memcpy(dstC0 + (ptrdiff_t) y * (dx / 2), s0.Cb().buf + (ptrdiff_t) y * s0.Cb().stride, sizeof(int16_t) * (dx / 2));
        memcpy(dstC1 + (ptrdiff_t) y * (dx / 2), s1.Cb().buf + (ptrdiff_t) y * s1.Cb().stride, sizeof(int16_t) * (dx / 2));
      }
      dstC0 += (dx / 2) * (dy / 2);
      dstC1 += (dx / 2) * (dy / 2);
    }
  }
  for (int l = 0; l < 2; l++) pred[l].destroy();
  delete slice;
  pics[0].destroy();
  pics[1].destroy();
  return 0;
}

// ---- symmetric MVD search: the reference's own InterSearch::xSymmetricMotionEstimation (InterSearch.cpp:4506-4518) ------
// planeCur / planeTar: luma planes of the reference pictures of eRefPicList (list 0 here) and of the other list (`margin`
// border samples on each side already extended); org: the original block.  Same field layout as vo_smvd_io.
struct RefSmvdIo
{
  int      x, y, w, h, picW, picH, maxCuW, maxCuH, bd, imv;
  int      curPredX, curPredY, tarPredX, tarPredY, curMvX, curMvY, tarMvX, tarMvY;
  int      clipBiPred, useHad, bcwIdx;
  double   lambda;
  uint64_t cost;
};
int ref_smvd_search(const int16_t* org, int orgStride, const int16_t* planeCur, const int16_t* planeTar, int planeStride, int margin,
                    int n, RefSmvdIo* ios)
{
  Probe&  p = probe();
  if (n <= 0) return 0;
  const int lumaW = ios[0].picW, lumaH = ios[0].picH;
  Picture pics[2];
  for (int l = 0; l < 2; l++)
  {
    Picture& pic = pics[l];
    pic.create(CHROMA_420, Size(lumaW, lumaH), 128, 128 + 16, false, 0);
    pic.unscaledPic = &pic;
    PelBuf reco = pic.getRecoBuf(COMPONENT_Y);
    if (margin > (int) pic.margin) return -1;
    const int16_t* plane = l ? planeTar : planeCur;
    for (int y = -margin; y < lumaH + margin; y++)
      memcpy(reco.buf + (ptrdiff_t) y * reco.stride - margin, plane + (ptrdiff_t) (y + margin) * planeStride,
             sizeof(int16_t) * (lumaW + 2 * margin));
  }
  if (p.m_tmpStorageLCU.bufs.empty())   // what InterSearch::init creates (InterSearch.cpp:254-258)
  {
    for (int i = 0; i < NUM_REF_PIC_LIST_01; i++) p.m_tmpPredStorage[i].create(UnitArea(CHROMA_420, Area(0, 0, MAX_CU_SIZE, MAX_CU_SIZE)));
    p.m_tmpStorageLCU.create(UnitArea(CHROMA_420, Area(0, 0, MAX_CU_SIZE, MAX_CU_SIZE)));
  }
  PPS pps;
  pps.setPicWidthInLumaSamples(lumaW);
  pps.setPicHeightInLumaSamples(lumaH);
  SPS sps;
  std::vector<uint64_t> shell((sizeof(CodingStructure) + 7) / 8, 0);
  CodingStructure* cs = reinterpret_cast<CodingStructure*>(shell.data());
  cs->sps = &sps;
  cs->pps = &pps;
  Slice* slice = new Slice();
  for (int l = 0; l < 2; l++)
  {
    slice->m_apcRefPicList[l][0] = &pics[l];
    slice->m_scalingRatio[l][0]  = SCALE_1X;
  }
  slice->setSPS(&sps);
  clipMv = clipMvInPic;
  p.cfg.setMCTSEncConstraint(false);
  PelStorage orgStore;
  orgStore.create(UnitArea(CHROMA_420, Area(0, 0, MAX_CU_SIZE, MAX_CU_SIZE)));
  for (int i = 0; i < n; i++)
  {
    RefSmvdIo& io = ios[i];
    sps.setBitDepth(CHANNEL_TYPE_LUMA, io.bd);
    sps.setBitDepth(CHANNEL_TYPE_CHROMA, io.bd);
    sps.setMaxCUWidth(io.maxCuW);
    sps.setMaxCUHeight(io.maxCuH);
    for (int c = 0; c < MAX_NUM_COMPONENT; c++) slice->m_clpRngs.comp[c] = makeClp(io.bd);
    slice->setDisableSATDForRD(!io.useHad);
    p.cfg.setClipForBiPredMeEnabled(io.clipBiPred != 0);
    p.rd.m_motionLambda = io.lambda;
    CodingUnit cu(CHROMA_420, Area(io.x, io.y, io.w, io.h));
    cu.imv    = io.imv;
    cu.affine = false;
    cu.BcwIdx = (uint8_t) io.bcwIdx;
    cu.slice  = slice;
    PredictionUnit pu(CHROMA_420, Area(io.x, io.y, io.w, io.h));
    pu.cu = &cu;
    pu.cs = cs;
    cs->slice = slice;
    const UnitArea rel(CHROMA_420, Area(0, 0, io.w, io.h));
    PelUnitBuf orgBuf = orgStore.getBuf(rel);
    for (int y = 0; y < io.h; y++) memcpy(orgBuf.Y().buf + (ptrdiff_t) y * orgBuf.Y().stride, org + (ptrdiff_t) (io.y + y) * orgStride + io.x, sizeof(int16_t) * io.w);
    Mv         curPred(io.curPredX, io.curPredY), tarPred(io.tarPredX, io.tarPredY);
    MvField    cur(Mv(io.curMvX, io.curMvY), 0), tar(Mv(io.tarMvX, io.tarMvY), 0);
    Distortion cost = io.cost;
    p.xSymmetricMotionEstimation(pu, orgBuf, curPred, tarPred, REF_PIC_LIST_0, cur, tar, cost, io.bcwIdx);
    io.curMvX = cur.mv.hor;
    io.curMvY = cur.mv.ver;
    io.tarMvX = tar.mv.hor;
    io.tarMvY = tar.mv.ver;
    io.cost   = cost;
  }
  orgStore.destroy();
  delete slice;
  pics[0].destroy();
  pics[1].destroy();
  return 0;
}

// AreaBuf<Pel>::removeHighFreq (Buffer.h:474-517): dst = 2*dst - src, optionally clipped (bi-pred ME target)
void ref_remove_high_freq(int16_t* dst, int dstStride, const int16_t* src, int srcStride, int w, int h, int clip, int bd)
{
  PelBuf  d(dst, dstStride, w, h);
  PelBuf  s(const_cast<int16_t*>(src), srcStride, w, h);
  d.removeHighFreq(s, clip != 0, makeClp(bd));
}

// AreaBuf<Pel>::addAvg (Buffer.cpp:467-507) through the SIMD table where the width allows
void ref_add_avg(const int16_t* s0, const int16_t* s1, int16_t* dst, int w, int h, int bd)
{
  PelBuf  d(dst, w, w, h);
  CPelBuf a(s0, w, w, h), b(s1, w, w, h);
  d.addAvg(a, b, makeClp(bd));
}

// AreaBuf<Pel>::addWeightedAvg (Buffer.cpp:365-396) / AreaBuf<T>::removeWeightHighFreq (Buffer.h:418-472): the BCW forms
void ref_add_weighted_avg(const int16_t* s0, const int16_t* s1, int16_t* dst, int w, int h, int bd, int bcwIdx)
{
  PelBuf  d(dst, w, w, h);
  CPelBuf a(s0, w, w, h), b(s1, w, w, h);
  d.addWeightedAvg(a, b, makeClp(bd), (int8_t) bcwIdx);
}
void ref_remove_weight_high_freq(int16_t* dst, const int16_t* src, int w, int h, int clip, int bd, int bcwWeight)
{
  PelBuf d(dst, w, w, h);
  PelBuf s(const_cast<int16_t*>(src), w, w, h);
  d.removeWeightHighFreq(s, clip != 0, makeClp(bd), (int8_t) bcwWeight);
}

// ---- InterSearch::xPatternSearchIntRefine (InterSearch.cpp:4172-4282), the reference's own member -----------------
// Same layout as vo_int_refine_io (oracle/vtm_oracle.h).
struct RefIntRefine
{
  int      imv;
  int      mvX, mvY;
  int      numCand;
  int      candX[2], candY[2];
  int      mvpIdx;
  uint32_t mvpIdxBits[2];
  uint32_t bits;
  double   fWeight;
  int      posX, posY, picW, picH, maxCuW, maxCuH;
  uint64_t cost;
};

void ref_int_refine(const RefSearchJob* j, RefIntRefine* io)
{
  Probe& p = probe();
  p.rd.m_motionLambda = j->lambdaMotion;
  p.cfg.setUseHADME(j->useHad != 0);
  p.cfg.setMCTSEncConstraint(false);
  p.m_lumaClpRng = makeClp(j->bitDepth);
  clipMv = clipMvInPic;   // EncGOP.cpp:2766

  PPS pps;
  pps.setPicWidthInLumaSamples(io->picW);
  pps.setPicHeightInLumaSamples(io->picH);
  SPS sps;
  sps.setMaxCUWidth(io->maxCuW);
  sps.setMaxCUHeight(io->maxCuH);
  sps.setWrapAroundEnabledFlag(false);
  Slice slice;
  slice.setDisableSATDForRD(false);
  std::vector<uint64_t> shell((sizeof(CodingStructure) + 7) / 8, 0);
  CodingStructure* cs = reinterpret_cast<CodingStructure*>(shell.data());
  cs->sps   = &sps;
  cs->pps   = &pps;
  cs->slice = &slice;
  CodingUnit cu(CHROMA_420, Area(io->posX, io->posY, j->w, j->h));
  cu.imv    = io->imv;
  cu.affine = false;
  PredictionUnit pu(CHROMA_420, Area(io->posX, io->posY, j->w, j->h));
  pu.cu = &cu;
  pu.cs = cs;

  CPelBuf pattern(j->org, j->orgStride, j->w, j->h);
  InterSearch::IntTZSearchStruct st;
  memset(&st, 0, sizeof(st));
  st.pcPatternKey = &pattern;
  st.piRefY       = j->refAtPU;
  st.iRefStride   = j->refStride;

  AMVPInfo amvp;
  amvp.numCand = io->numCand;
  for (int i = 0; i < 2; i++) amvp.mvCand[i] = Mv(io->candX[i], io->candY[i]);
  for (int i = 0; i < 2; i++) p.m_auiMVPIdxCost[i][AMVP_MAX_NUM_CANDS] = io->mvpIdxBits[i];

  Mv         mv(io->mvX, io->mvY);
  Mv         pred = amvp.mvCand[io->mvpIdx];
  int        idx  = io->mvpIdx;
  uint32_t   bits = io->bits;
  Distortion cost = 0;
  p.xPatternSearchIntRefine(pu, st, mv, pred, idx, bits, cost, amvp, io->fWeight);
  io->mvX    = mv.hor;
  io->mvY    = mv.ver;
  io->mvpIdx = idx;
  io->bits   = bits;
  io->cost   = cost;
}

// ---- InterSearch::xTZSearch (InterSearch.cpp:3640-3974), the reference's own member ---------------------------------
// Same layout as vo_tz_params (oracle/vtm_oracle.h).
struct RefTzParams
{
  int startX, startY;
  int hasInt2Nx2N, int2Nx2NX, int2Nx2NY;
  int nSeeds;
  int seedX[16], seedY[16];
  int searchRange;
  int extended, fast, firstSearchStop;
  int posX, posY, picW, picH, maxCuW, maxCuH;
  int selective;   // xTZSearchSelective (MESEARCH_SELECTIVE) instead of xTZSearch
};

// per-thread environment of the xTZSearch calls: parameter sets and the CodingStructure shell are built once
struct TzEnv
{
  PPS                       pps;
  SPS                       sps;
  std::vector<uint64_t>     shell;
  std::vector<BlkUniMvInfo> list;
  CodingStructure*          cs;
  TzEnv() : shell((sizeof(CodingStructure) + 7) / 8, 0), list(15)
  {
    sps.setWrapAroundEnabledFlag(false);
    // xClipMv (InterSearch.cpp:7735-7764) looks the sub-picture up: one sub-picture, not treated as a picture
    pps.m_numSubPics = 1;
    pps.m_subPics.resize(1);
    pps.m_subPics[0].setTreatedAsPicFlag(false);
    pps.setWrapAroundEnabledFlag(false);
    cs      = reinterpret_cast<CodingStructure*>(shell.data());
    cs->sps = &sps;
    cs->pps = &pps;
  }
};
thread_local TzEnv* t_tzEnv = nullptr;

void ref_tz_search(const RefSearchJob* j, const RefTzParams* t, int* mvx, int* mvy, uint64_t* sad)
{
  Probe& p = probe();
  if (!t_tzEnv) t_tzEnv = new TzEnv();
  TzEnv& e = *t_tzEnv;
  p.rd.m_motionLambda = j->lambdaMotion;
  p.rd.setPredictor(Mv(j->predQx, j->predQy));
  p.rd.setCostScale(2);
  p.cfg.setMCTSEncConstraint(false);
  p.cfg.setFastMEAssumingSmootherMVEnabled(t->firstSearchStop != 0);
  p.cfg.setUseHashME(false);
  p.m_lumaClpRng   = makeClp(j->bitDepth);
  p.m_iSearchRange = t->searchRange;
  clipMv = clipMvInPic;

  e.pps.setPicWidthInLumaSamples(t->picW);
  e.pps.setPicHeightInLumaSamples(t->picH);
  e.sps.setMaxCUWidth(t->maxCuW);
  e.sps.setMaxCUHeight(t->maxCuH);
  CodingUnit cu(CHROMA_420, Area(t->posX, t->posY, j->w, j->h));
  cu.imv    = 0;
  cu.affine = false;
  PredictionUnit pu(CHROMA_420, Area(t->posX, t->posY, j->w, j->h));
  pu.cu = &cu;
  pu.cs = e.cs;

  // history of uni-directional MVs (m_uniMvList): entry i of the caller's list is the i-th newest
  const int maxSize = 15;
  p.m_uniMvList        = e.list.data();
  p.m_uniMvListMaxSize = maxSize;
  p.m_uniMvListSize    = t->nSeeds;
  p.m_uniMvListIdx     = t->nSeeds % maxSize;
  for (int i = 0; i < t->nSeeds; i++)
    e.list[(p.m_uniMvListIdx - 1 - i + maxSize) % maxSize].uniMvs[0][0] = Mv(t->seedX[i], t->seedY[i]);

  CPelBuf pattern(j->org, j->orgStride, j->w, j->h);
  InterSearch::IntTZSearchStruct st;
  memset(&st, 0, sizeof(st));
  st.pcPatternKey = &pattern;
  st.piRefY       = j->refAtPU;
  st.iRefStride   = j->refStride;
  st.imvShift     = j->imvShift;
  st.subShiftMode = j->subShiftMode;

  Mv         mv(t->startX, t->startY);
  Mv         int2Nx2N(t->int2Nx2NX, t->int2Nx2NY);
  Distortion cost = 0;
  if (t->selective)
    p.xTZSearchSelective(pu, REF_PIC_LIST_0, 0, st, mv, cost, t->hasInt2Nx2N ? &int2Nx2N : nullptr);
  else
    p.xTZSearch(pu, REF_PIC_LIST_0, 0, st, mv, cost, t->hasInt2Nx2N ? &int2Nx2N : nullptr, t->extended != 0, t->fast != 0);
  p.m_uniMvList = nullptr;
  *mvx = mv.hor;
  *mvy = mv.ver;
  *sad = cost;
}

// n TZ searches over nThreads workers; mv = n x {x, y}; returns wall seconds (CPU baseline of the batched TZ search)
double ref_tz_batch(const RefSearchJob* jobs, const RefTzParams* tz, int n, int nThreads, int* mv, uint64_t* sad)
{
  if (nThreads < 1) nThreads = 1;
  std::atomic<int> next(0);
  auto t0 = std::chrono::steady_clock::now();
  auto worker = [&]() {
    for (;;)
    {
      int i = next.fetch_add(16);
      if (i >= n) break;
      const int e = i + 16 < n ? i + 16 : n;
      for (; i < e; i++) ref_tz_search(&jobs[i], &tz[i], &mv[2 * i], &mv[2 * i + 1], &sad[i]);
    }
  };
  if (nThreads == 1)
    worker();
  else
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nThreads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- EncTemporalFilter::motionEstimation (EncTemporalFilter.cpp:448-466), the reference's own member ----------------
// org / ref: sample (0,0) of planes carrying `margin` (>= 0) valid border samples; the shim copies the picture area
// into PelStorage buffers with the filter's own padding (128) and extends the border as EncTemporalFilter::filter does
// (:169-204).  mv: (width/4) x (height/4) entries {x, y, error}.  Returns the seconds spent in motionEstimation.
double ref_mctf_me(const int16_t* org, int orgStride, const int16_t* ref, int refStride, int width, int height, int bitDepth,
                   int32_t* mv)
{
  EncTemporalFilter tf;
  tf.m_chromaFormatIDC = CHROMA_400;
  tf.m_sourceWidth     = width;
  tf.m_sourceHeight    = height;
  for (int i = 0; i < MAX_NUM_CHANNEL_TYPE; i++) tf.m_internalBitDepth[i] = bitDepth;
  const int  pad = EncTemporalFilter::m_padding;
  const Area area(0, 0, width, height);
  PelStorage o, b, o2, o4;
  o.create(CHROMA_400, area, 0, pad);
  b.create(CHROMA_400, area, 0, pad);
  for (int y = 0; y < height; y++)
  {
    memcpy(o.Y().buf + (ptrdiff_t) y * o.Y().stride, org + (ptrdiff_t) y * orgStride, sizeof(int16_t) * width);
    memcpy(b.Y().buf + (ptrdiff_t) y * b.Y().stride, ref + (ptrdiff_t) y * refStride, sizeof(int16_t) * width);
  }
  o.extendBorderPel(pad, pad);
  b.extendBorderPel(pad, pad);
  tf.subsampleLuma(o, o2);
  tf.subsampleLuma(o2, o4);
  Array2D<MotionVector> mvs;
  mvs.allocate(width / 4, height / 4);
  auto t0 = std::chrono::steady_clock::now();
  tf.motionEstimation(mvs, o, b, o2, o4);
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (int y = 0; y < height / 4; y++)
    for (int x = 0; x < width / 4; x++)
    {
      const MotionVector& m = mvs.get(x, y);
      int32_t*            d = mv + 3 * ((size_t) y * (width / 4) + x);
      d[0] = m.x;
      d[1] = m.y;
      d[2] = m.error;
    }
  return sec;
}

// ---- EncTemporalFilter::applyMotion (EncTemporalFilter.cpp:470-552), the reference's own member ---------------------
// luma: lumaW x lumaH plane, chroma: (lumaW/2) x (lumaH/2) plane used for Cb and Cr (4:2:0); both without border (the
// shim pads them like EncTemporalFilter::filter).  mv: the (lumaW/4) x (lumaH/4) field of ref_mctf_me.
// dstY / dstC: packed outputs (row stride = component width); samples the reference does not write are zero.
void ref_mctf_apply_motion(const int16_t* luma, const int16_t* chroma, int lumaW, int lumaH, int bitDepth, const int32_t* mv,
                           int16_t* dstY, int16_t* dstC)
{
  EncTemporalFilter tf;
  tf.m_chromaFormatIDC = CHROMA_420;
  tf.m_sourceWidth     = lumaW;
  tf.m_sourceHeight    = lumaH;
  for (int i = 0; i < MAX_NUM_CHANNEL_TYPE; i++) tf.m_internalBitDepth[i] = bitDepth;
  const int  pad = EncTemporalFilter::m_padding;
  const Area area(0, 0, lumaW, lumaH);
  PelStorage in, out;
  in.create(CHROMA_420, area, 0, pad);
  out.create(CHROMA_420, area, 0, pad);
  for (int c = 0; c < 3; c++)
  {
    const int      w = c ? lumaW / 2 : lumaW, h = c ? lumaH / 2 : lumaH;
    const int16_t* s = c ? chroma : luma;
    for (int y = 0; y < h; y++)
    {
      memcpy(in.bufs[c].buf + (ptrdiff_t) y * in.bufs[c].stride, s + (ptrdiff_t) y * w, sizeof(int16_t) * w);
      memset(out.bufs[c].buf + (ptrdiff_t) y * out.bufs[c].stride, 0, sizeof(int16_t) * w);
    }
  }
  in.extendBorderPel(pad, pad);
  Array2D<MotionVector> mvs;
  mvs.allocate(lumaW / 4, lumaH / 4);
  for (int y = 0; y < lumaH / 4; y++)
    for (int x = 0; x < lumaW / 4; x++)
    {
      const int32_t* m = mv + 3 * ((size_t) y * (lumaW / 4) + x);
      mvs.get(x, y).set(m[0], m[1], m[2]);
    }
  tf.applyMotion(mvs, in, out);
  for (int c = 0; c < 2; c++)
  {
    const int w = c ? lumaW / 2 : lumaW, h = c ? lumaH / 2 : lumaH;
    int16_t*  d = c ? dstC : dstY;
    for (int y = 0; y < h; y++) memcpy(d + (ptrdiff_t) y * w, out.bufs[c].buf + (ptrdiff_t) y * out.bufs[c].stride, sizeof(int16_t) * w);
  }
}

// ---- EncTemporalFilter::bilateralFilter (EncTemporalFilter.cpp:555-623), the reference's own member ------------------
// org / neighbours: luma (lumaW x lumaH) and chroma ((lumaW/2) x (lumaH/2), used for Cb and Cr) planes without border; mv:
// numRefs vector fields of (lumaW/4) x (lumaH/4) entries {x, y, error} (ref_mctf_me); offsets: POC distances.
// dstY / dstC: packed filtered planes (Cb).
void ref_mctf_bilateral(const int16_t* orgY, const int16_t* orgC, const int16_t* const* refY, const int16_t* const* refC, int numRefs,
                        const int32_t* mv, const int* offsets, int lumaW, int lumaH, int bitDepth, int qp, double overallStrength,
                        int16_t* dstY, int16_t* dstC)
{
  EncTemporalFilter tf;
  tf.m_chromaFormatIDC = CHROMA_420;
  tf.m_sourceWidth     = lumaW;
  tf.m_sourceHeight    = lumaH;
  tf.m_QP              = qp;
  tf.m_area            = Area(0, 0, lumaW, lumaH);
  for (int i = 0; i < MAX_NUM_CHANNEL_TYPE; i++) tf.m_internalBitDepth[i] = bitDepth;
  const int  pad = EncTemporalFilter::m_padding;
  const Area area(0, 0, lumaW, lumaH);
  auto load = [&](PelStorage& st, const int16_t* y, const int16_t* c) {
    st.create(CHROMA_420, area, 0, pad);
    for (int k = 0; k < 3; k++)
    {
      const int      w = k ? lumaW / 2 : lumaW, h = k ? lumaH / 2 : lumaH;
      const int16_t* s = k ? c : y;
      for (int r = 0; r < h; r++) memcpy(st.bufs[k].buf + (ptrdiff_t) r * st.bufs[k].stride, s + (ptrdiff_t) r * w, sizeof(int16_t) * w);
    }
    st.extendBorderPel(pad, pad);
  };
  PelStorage org, out;
  load(org, orgY, orgC);
  out.create(CHROMA_420, area, 0, pad);
  std::deque<TemporalFilterSourcePicInfo> infos(numRefs);
  const size_t fieldSize = (size_t) (lumaW / 4) * (lumaH / 4) * 3;
  for (int i = 0; i < numRefs; i++)
  {
    load(infos[i].picBuffer, refY[i], refC[i]);
    infos[i].origOffset = offsets[i];
    infos[i].mvs.allocate(lumaW / 4, lumaH / 4);
    for (int y = 0; y < lumaH / 4; y++)
      for (int x = 0; x < lumaW / 4; x++)
      {
        const int32_t* m = mv + i * fieldSize + 3 * ((size_t) y * (lumaW / 4) + x);
        infos[i].mvs.get(x, y).set(m[0], m[1], m[2]);
      }
  }
  tf.bilateralFilter(org, infos, out, overallStrength);
  for (int c = 0; c < 2; c++)
  {
    const int w = c ? lumaW / 2 : lumaW, h = c ? lumaH / 2 : lumaH;
    int16_t*  d = c ? dstC : dstY;
    for (int y = 0; y < h; y++) memcpy(d + (ptrdiff_t) y * w, out.bufs[c].buf + (ptrdiff_t) y * out.bufs[c].stride, sizeof(int16_t) * w);
  }
}

// Batch driver used as the CPU baseline: nThreads workers over disjoint job ranges.
// Returns wall seconds spent in the searches (steady_clock around the work only).
double ref_search_batch(const RefSearchJob* jobs, RefSearchResult* res, int n, int nThreads)
{
  if (nThreads < 1) nThreads = 1;
  std::atomic<int> next(0);
  auto t0 = std::chrono::steady_clock::now();
  auto worker = [&]() {
    for (;;)
    {
      int i = next.fetch_add(16);
      if (i >= n) break;
      int e = i + 16 < n ? i + 16 : n;
      for (; i < e; i++) ref_search(&jobs[i], &res[i]);
    }
  };
  if (nThreads == 1)
  {
    worker();
  }
  else
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nThreads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// AffineGradientSearch's own table entries (SIMD where the reference installs them): Sobel filters and xEqualCoeffComputer
void ref_affine_sobel(int vertical, const int16_t* pred, int predStride, int* deriv, int derivStride, int w, int h)
{
  static AffineGradientSearch ags;
  (vertical ? ags.m_VerticalSobelFilter : ags.m_HorizontalSobelFilter)(const_cast<Pel*>(pred), predStride, deriv, derivStride, w, h);
}

void ref_affine_equal_coeff(const int16_t* residue, int residueStride, int* d0, int* d1, int derivStride, int64_t* coeff, int w, int h,
                            int sixParam)
{
  static AffineGradientSearch ags;
  int* dd[2] = { d0, d1 };
  ags.m_EqualCoeffComputer(const_cast<Pel*>(residue), residueStride, dd, derivStride, reinterpret_cast<int64_t(*)[7]>(coeff), w, h, sixParam != 0);
}

}   // extern "C"
