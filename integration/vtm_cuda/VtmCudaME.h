// VTM-side glue of libvtmme (this repository's own code; added to a scratch copy of VTM 9.3 by
// integration/apply_patch.py under source/Lib/CommonLib/cuda/).  Mirrors how x86/InitX86.cpp hooks the SIMD kernels:
//   RdCost::initRdCostCUDA(), InterpolationFilter::initInterpolationFilterCUDA()  (cuda/InitCUDA.cpp)
//   vtmcuda::search()  — the batched entry used by InterSearch::xMotionEstimation in place of
//                        xPatternSearch (InterSearch.cpp:3432) + xPatternSearchFracDIF (:3476)
#pragma once
#include <cstdint>

class Picture;

namespace vtmcuda
{
// VTMME_ENABLE=1 in the environment switches the motion search to the GPU (read once; cf. --SIMD= in encmain.cpp:92-100)
bool enabled();
// VTMME_TABLE_HOOKS=1 additionally routes the distortion / interpolation dispatch tables through the GPU
bool tableHooksEnabled();

struct SearchIn
{
  const Picture* refPic;        // reference picture (recon plane, border extended)
  int            x, y, w, h;    // PU luma rectangle
  const int16_t* org;           // pattern (cStruct.pcPatternKey), host memory
  int            orgStride;
  int            srLeft, srRight, srTop, srBottom;   // cStruct.searchRange
  int            predQx, predQy;                     // predictor, quarter-pel
  int            imvShift;
  int            subShiftMode;  // cStruct.subShiftMode (0 or 2; 1 for the TZ searches of MESEARCH_SELECTIVE)
  int            bitDepth;
  bool           useHad, useAltHpel;
  int            doFrac;        // 0 integer search only, 1 + xPatternSearchFracDIF, 2 + xPatternSearchIntRefine (AMVR)
  double         lambdaMotion;
  // doFrac == 2: what xPatternSearchIntRefine receives (InterSearch.cpp:4172), MVs in MV_PRECISION_INTERNAL
  int            imv, numCand, candX[2], candY[2], mvpIdx;
  uint32_t       mvpIdxBits[2], bits;
  int            picW, picH, maxCuW, maxCuH;
  double         fWeight;
  // tzSearch: the integer search is xTZSearch (FastSearch=1/3) or xTZSearchSelective (FastSearch=2, tzSelective) instead
  // of the full search; sr* are ignored
  bool           tzSearch, tzExtended, tzFast, tzFirstSearchStop, tzSelective;
  int            tzStartX, tzStartY;       // rcMv on entry, MV_PRECISION_INTERNAL
  int            tzSearchRange;            // m_iSearchRange
  int            tzNumSeeds, tzSeedX[16], tzSeedY[16];   // m_uniMvList entries of (list, ref), newest first
};

struct SearchOut
{
  int      mvX, mvY;
  uint64_t intSad;
  int      halfX, halfY, qterX, qterY;
  uint64_t fracCost;
  int      amvrMvX, amvrMvY, mvpIdx;   // doFrac == 2: rcMv (internal precision), riMVPIdx,
  uint32_t bits;                       //   ruiBits,
  uint64_t cost;                       //   ruiCost
};

// Per-PU batching: InterSearch::predInterSearch first walks its (list, reference) loop in *collect* mode — xMotionEstimation runs
// up to the point where it would search and hands the search's inputs to collect mode instead — then endCollect() runs all
// collected searches in ONE vtmme_search call.  The real loop that follows asks search() as before; a request whose inputs are
// identical, field by field, to a collected one is answered from that batch (anything else is searched on its own), so the
// results cannot depend on the batching.  VTMME_BATCH=0 disables it.
bool batching();
bool collecting();
void beginCollect();
void endCollect();

// Runs one xMotionEstimation search on the GPU.  Any failure of the CUDA path is fatal (THROW): there is no
// CPU fallback once the GPU path is enabled.
void search( const SearchIn& in, SearchOut& out );

// GOP-based temporal filter (EncTemporalFilter::filter, EncoderLib/EncTemporalFilter.cpp:133-236) on the GPU: the motion
// estimation of every neighbouring picture against the original (one batched vtmme_mctf_me call), applyMotion and the
// bilateral weighting of each component.  Planes are plain pointers (sample (0,0), row stride in samples); the library
// re-makes the replicated border the filter pads its pictures with.  weights: per component numRefs tables of
// (1 << bitDepth[c]) doubles, computed by the caller with its own exp() (EncTemporalFilter.cpp:595-610), so the filtered
// samples are the reference's bit for bit.
struct TfPlane
{
  const int16_t* buf;
  int            stride, width, height;
};
void temporalFilter( int numRefs, int numComp, const TfPlane* org /*[numComp]*/, const TfPlane* refs /*[numRefs][numComp]*/, int csx, int csy,
                     const int* bitDepth /*[numComp]*/, const double* const* weights /*[numComp]*/, int16_t* const* dst /*[numComp]*/,
                     const int* dstStride );

// per-block table entries (used by InitCUDA.cpp)
uint64_t distHost( int kind, const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w, int h, int subShift );
void     interpHost( int comp, int vertical, const int16_t* src, int srcStride, int16_t* dst, int dstStride, int w, int h,
                     int frac, int isFirst, int isLast, int bitDepth, int useAltHpel );
// table-entry flavour of the filters: explicit taps (what m_filterHor/m_filterVer/m_filterCopy entries receive)
void     filterHost( int taps, int vertical, int isFirst, int isLast, int copy, const int16_t* src, int srcStride, int16_t* dst,
                     int dstStride, int w, int h, const int16_t* coeff, int bitDepth );
// AffineGradientSearch's table entries (InitCUDA.cpp)
void     affineSobelHost( int vertical, const int16_t* pred, int predStride, int w, int h, int* deriv, int derivStride );
void     affineEqualCoeffHost( const int16_t* residue, int residueStride, const int* d0, const int* d1, int derivStride, int w, int h, int sixParam,
                               int64_t* coeff );
void     printStats();
}   // namespace vtmcuda
