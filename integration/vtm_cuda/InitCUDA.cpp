// RdCost::initRdCostCUDA() / InterpolationFilter::initInterpolationFilterCUDA(): the CUDA siblings of
// initRdCostX86() (x86/InitX86.cpp:104-122) and initInterpolationFilterX86() (x86/InitX86.cpp:57-75).
// They override the same dispatch-table entries with the same signatures; an entry delegates to the entry it
// replaced for what the X86 versions also leave to the scalar code (weighted prediction, bit depth > 10, MR-SAD,
// masks, DMVR bilinear mode — x86/RdCostX86.h:213,344,2157).
//
// One block per call with host pointers is launch-bound by construction (SURVEY.md §7 "hard parts" 3): these
// hooks exist for table-level parity and the micro-benchmark; encoder throughput comes from vtmcuda::search().
// They are only installed when VTMME_TABLE_HOOKS=1.
#include "CommonLib/AffineGradientSearch.h"
#include "CommonLib/CommonDef.h"
#include "CommonLib/InterpolationFilter.h"
#include "CommonLib/RdCost.h"
#include "VtmCudaME.h"

#include <cstdio>
#include <cstdlib>

namespace
{
FpDistFunc s_prevDist[DF_TOTAL_FUNCTIONS];
// calls answered by the GPU / handed to the entry that was replaced; printed at exit (VTMME_TABLE_HOOKS=2: count
// only — every call is delegated, which sizes a run without a GPU)
unsigned long long s_distGpu = 0, s_distPrev = 0, s_filtGpu = 0, s_filtPrev = 0, s_affGpu = 0, s_affPrev = 0;
bool               s_countOnly = false;

void printHookStats()
{
  fprintf( stderr, "[vtmcuda] table hooks: distortion %llu on the GPU / %llu delegated, filters %llu on the GPU / %llu delegated, "
                   "affine gradient entries %llu on the GPU / %llu delegated%s\n",
           s_distGpu, s_distPrev, s_filtGpu, s_filtPrev, s_affGpu, s_affPrev, s_countOnly ? " (count only)" : "" );
}

template<int DF> Distortion distCuda( const DistParam& dp )
{
  const int w = dp.org.width, h = dp.org.height;
  if( dp.applyWeight || dp.bitDepth > 10 || dp.useMR || w < 2 || h < 2 || w > 128 || h > 128 || ( w & 1 ) || ( h & 1 ) || dp.step != 1 )
  {
    s_distPrev++;
    return s_prevDist[DF]( dp );
  }
  const bool had = DF >= DF_HAD && DF <= DF_HAD16N;
  if( ( !had && ( h >> dp.subShift ) < 1 ) || s_countOnly )
  {
    s_distPrev++;
    return s_prevDist[DF]( dp );
  }
  s_distGpu++;
  return vtmcuda::distHost( had ? 1 : 0, dp.org.buf, dp.org.stride, dp.cur.buf, dp.cur.stride, w, h, had ? 0 : dp.subShift );
}

typedef void ( *FilterFn )( const ClpRng&, Pel const*, int, Pel*, int, int, int, TFilterCoeff const*, bool );
typedef void ( *CopyFn )( const ClpRng&, Pel const*, int, Pel*, int, int, int, bool );
FilterFn s_prevHor[3][2][2], s_prevVer[3][2][2];
CopyFn   s_prevCopy[2][2];

template<int T, bool VER, bool FIRST, bool LAST>
void filterCuda( const ClpRng& clpRng, Pel const* src, int srcStride, Pel* dst, int dstStride, int width, int height,
                 TFilterCoeff const* coeff, bool biMCForDMVR )
{
  if( biMCForDMVR || clpRng.bd > 10 || clpRng.bd < 8 || width > 256 || height > 256 || clpRng.min != 0 || clpRng.max != ( 1 << clpRng.bd ) - 1 || s_countOnly )
  {
    s_filtPrev++;
    ( VER ? s_prevVer : s_prevHor )[T][FIRST][LAST]( clpRng, src, srcStride, dst, dstStride, width, height, coeff, biMCForDMVR );
    return;
  }
  s_filtGpu++;
  const int taps = T == 0 ? 8 : ( T == 1 ? 4 : 2 );
  int16_t   c[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
  for( int k = 0; k < taps; k++ ) c[k] = coeff[k];
  vtmcuda::filterHost( taps, VER, FIRST, LAST, 0, src, srcStride, dst, dstStride, width, height, c, clpRng.bd );
}

template<bool FIRST, bool LAST>
void copyCuda( const ClpRng& clpRng, Pel const* src, int srcStride, Pel* dst, int dstStride, int width, int height, bool biMCForDMVR )
{
  if( biMCForDMVR || clpRng.bd > 10 || clpRng.bd < 8 || width > 256 || height > 256 || clpRng.min != 0 || clpRng.max != ( 1 << clpRng.bd ) - 1 || s_countOnly )
  {
    s_filtPrev++;
    s_prevCopy[FIRST][LAST]( clpRng, src, srcStride, dst, dstStride, width, height, biMCForDMVR );
    return;
  }
  s_filtGpu++;
  vtmcuda::filterHost( 8, 0, FIRST, LAST, 1, src, srcStride, dst, dstStride, width, height, nullptr, clpRng.bd );
}
}   // namespace

void RdCost::initRdCostCUDA()
{
  if( !vtmcuda::tableHooksEnabled() ) return;
  static bool done = false;   // the table is static and every RdCost constructor runs init(): stay idempotent
  if( !done )
  {
    for( int i = 0; i < DF_TOTAL_FUNCTIONS; i++ ) s_prevDist[i] = m_afpDistortFunc[i];
    done = true;
    const char* e = getenv( "VTMME_TABLE_HOOKS" );
    s_countOnly   = e && e[0] == '2';
    atexit( printHookStats );
  }
#define VTMCUDA_DIST( DF ) m_afpDistortFunc[DF] = distCuda<DF>;
  VTMCUDA_DIST( DF_SAD ) VTMCUDA_DIST( DF_SAD2 ) VTMCUDA_DIST( DF_SAD4 ) VTMCUDA_DIST( DF_SAD8 ) VTMCUDA_DIST( DF_SAD16 )
  VTMCUDA_DIST( DF_SAD32 ) VTMCUDA_DIST( DF_SAD64 ) VTMCUDA_DIST( DF_SAD16N ) VTMCUDA_DIST( DF_SAD12 ) VTMCUDA_DIST( DF_SAD24 )
  VTMCUDA_DIST( DF_SAD48 )
  VTMCUDA_DIST( DF_HAD ) VTMCUDA_DIST( DF_HAD2 ) VTMCUDA_DIST( DF_HAD4 ) VTMCUDA_DIST( DF_HAD8 ) VTMCUDA_DIST( DF_HAD16 )
  VTMCUDA_DIST( DF_HAD32 ) VTMCUDA_DIST( DF_HAD64 ) VTMCUDA_DIST( DF_HAD16N )
#undef VTMCUDA_DIST
}

void InterpolationFilter::initInterpolationFilterCUDA()
{
  if( !vtmcuda::tableHooksEnabled() ) return;
  for( int t = 0; t < 3; t++ )
    for( int f = 0; f < 2; f++ )
      for( int l = 0; l < 2; l++ )
      {
        s_prevHor[t][f][l] = m_filterHor[t][f][l];
        s_prevVer[t][f][l] = m_filterVer[t][f][l];
      }
  for( int f = 0; f < 2; f++ )
    for( int l = 0; l < 2; l++ ) s_prevCopy[f][l] = m_filterCopy[f][l];
  // luma 8-tap and chroma 4-tap, first/last stage combinations used by the codec (InterpolationFilter.cpp:335-373)
  m_filterHor[0][1][0] = filterCuda<0, false, true, false>;
  m_filterHor[0][1][1] = filterCuda<0, false, true, true>;
  m_filterHor[1][1][0] = filterCuda<1, false, true, false>;
  m_filterHor[1][1][1] = filterCuda<1, false, true, true>;
  m_filterVer[0][0][0] = filterCuda<0, true, false, false>;
  m_filterVer[0][0][1] = filterCuda<0, true, false, true>;
  m_filterVer[0][1][0] = filterCuda<0, true, true, false>;
  m_filterVer[0][1][1] = filterCuda<0, true, true, true>;
  m_filterVer[1][0][0] = filterCuda<1, true, false, false>;
  m_filterVer[1][0][1] = filterCuda<1, true, false, true>;
  m_filterVer[1][1][0] = filterCuda<1, true, true, false>;
  m_filterVer[1][1][1] = filterCuda<1, true, true, true>;
  m_filterCopy[0][1]   = copyCuda<false, true>;
  m_filterCopy[1][0]   = copyCuda<true, false>;
}

// AffineGradientSearch::initAffineGradientSearchCUDA(): the CUDA sibling of initAffineGradientSearchX86()
// (x86/AffineGradientSearchX86.h:311-317) — the Sobel filters and xEqualCoeffComputer of the affine motion estimation.
namespace
{
typedef void ( *SobelFn )( Pel* const, const int, int* const, const int, const int, const int );
typedef void ( *EqualFn )( Pel*, int, int**, int, int64_t ( * )[7], int, int, bool );
SobelFn s_prevSobelH = nullptr, s_prevSobelV = nullptr;
EqualFn s_prevEqual  = nullptr;

bool affineOnGpu( int w, int h ) { return !s_countOnly && w >= 4 && h >= 4 && w <= 128 && h <= 128; }

void sobelHCuda( Pel* const pPred, const int predStride, int* const pDerivate, const int derivateBufStride, const int width, const int height )
{
  if( !affineOnGpu( width, height ) ) { s_affPrev++; s_prevSobelH( pPred, predStride, pDerivate, derivateBufStride, width, height ); return; }
  s_affGpu++;
  vtmcuda::affineSobelHost( 0, pPred, predStride, width, height, pDerivate, derivateBufStride );
}

void sobelVCuda( Pel* const pPred, const int predStride, int* const pDerivate, const int derivateBufStride, const int width, const int height )
{
  if( !affineOnGpu( width, height ) ) { s_affPrev++; s_prevSobelV( pPred, predStride, pDerivate, derivateBufStride, width, height ); return; }
  s_affGpu++;
  vtmcuda::affineSobelHost( 1, pPred, predStride, width, height, pDerivate, derivateBufStride );
}

void equalCoeffCuda( Pel* pResidue, int residueStride, int** ppDerivate, int derivateBufStride, int64_t ( *pEqualCoeff )[7], int width, int height, bool b6Param )
{
  if( !affineOnGpu( width, height ) ) { s_affPrev++; s_prevEqual( pResidue, residueStride, ppDerivate, derivateBufStride, pEqualCoeff, width, height, b6Param ); return; }
  s_affGpu++;
  vtmcuda::affineEqualCoeffHost( pResidue, residueStride, ppDerivate[0], ppDerivate[1], derivateBufStride, width, height, b6Param ? 1 : 0, &pEqualCoeff[0][0] );
}
}   // namespace

void AffineGradientSearch::initAffineGradientSearchCUDA()
{
  if( !vtmcuda::tableHooksEnabled() ) return;
  if( !s_prevEqual )   // the entries every instance starts from are the same: remember them once
  {
    s_prevSobelH = m_HorizontalSobelFilter;
    s_prevSobelV = m_VerticalSobelFilter;
    s_prevEqual  = m_EqualCoeffComputer;
  }
  m_HorizontalSobelFilter = sobelHCuda;
  m_VerticalSobelFilter   = sobelVCuda;
  m_EqualCoeffComputer    = equalCoeffCuda;
}
