#!/usr/bin/env python
"""Creates a patched scratch copy of VTM 9.3 with the libvtmme hooks.

  python integration/apply_patch.py [/root/reference] [oracle/_ref/patched]

Nothing of the reference is stored in this repository: the script copies <ref>/source into the scratch directory
(git-ignored) and inserts a handful of lines at anchors; the new files it adds (cuda/VtmCudaME.*, cuda/InitCUDA.cpp)
are this repository's own (integration/vtm_cuda/).  INTEGRATION.md lists every insertion.
"""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def edit(path, fn):
    src = open(path).read()
    out = fn(src)
    assert out != src, "anchor not found in " + path
    open(path, "w").write(out)


def once(src, anchor, repl, count=1):
    assert src.count(anchor) >= 1, anchor
    return src.replace(anchor, repl, count)


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "oracle", "_ref", "patched")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(os.path.join(ref, "source"), os.path.join(dst, "source"))
    lib = os.path.join(dst, "source", "Lib")
    cuda = os.path.join(lib, "CommonLib", "cuda")
    os.makedirs(cuda)
    for f in ("VtmCudaME.h", "VtmCudaME.cpp", "InitCUDA.cpp"):
        shutil.copy(os.path.join(HERE, "vtm_cuda", f), cuda)

    # 1. RdCost.h: declare the hook next to initRdCostX86(), expose the selected motion lambda
    edit(os.path.join(lib, "CommonLib", "RdCost.h"), lambda s: once(
        s, "  void           setDistParam( DistParam &rcDP, const CPelBuf &org, const Pel* piRefY , int iRefStride, int bitDepth,",
        "  void          initRdCostCUDA();                                              // libvtmme\n"
        "  double        getSelectedMotionLambda() const { return m_motionLambda; }     // libvtmme\n"
        "  void           setDistParam( DistParam &rcDP, const CPelBuf &org, const Pel* piRefY , int iRefStride, int bitDepth,"))
    # 2. RdCost.cpp: call it right after initRdCostX86() in RdCost::init()
    edit(os.path.join(lib, "CommonLib", "RdCost.cpp"), lambda s: once(
        s, "  m_costMode                   = COST_STANDARD_LOSSY;",
        "  initRdCostCUDA();   // libvtmme: after the X86 table, same entries\n\n  m_costMode                   = COST_STANDARD_LOSSY;"))
    # 3. InterpolationFilter.h / .cpp
    edit(os.path.join(lib, "CommonLib", "InterpolationFilter.h"), lambda s: once(
        s, "  void filterHor(const ComponentID compID, Pel const* src,",
        "  void initInterpolationFilterCUDA();   // libvtmme\n  void filterHor(const ComponentID compID, Pel const* src,"))
    edit(os.path.join(lib, "CommonLib", "InterpolationFilter.cpp"), lambda s: re.sub(
        r"(void InterpolationFilter::initInterpolationFilter\( bool enable \)\n\{\n(?:.*\n)*?#endif\n#endif\n)\}",
        r"\1  if ( enable )\n  {\n    initInterpolationFilterCUDA();   // libvtmme\n  }\n}", s, count=1))

    # 3b. AffineGradientSearch: the third dispatch table (Sobel filters and xEqualCoeffComputer of the affine ME)
    edit(os.path.join(lib, "CommonLib", "AffineGradientSearch.h"), lambda s: once(
        s, "  void initAffineGradientSearchX86();", "  void initAffineGradientSearchX86();\n  void initAffineGradientSearchCUDA();   // libvtmme"))
    edit(os.path.join(lib, "CommonLib", "AffineGradientSearch.cpp"), lambda s: re.sub(
        r"(AffineGradientSearch::AffineGradientSearch\(\)\n\{\n(?:.*\n)*?#endif\n#endif\n)\}",
        r"\1  initAffineGradientSearchCUDA();   // libvtmme: after the X86 entries\n}", s, count=1))

    # 4. InterSearch.cpp: xMotionEstimation calls the GPU entry instead of xPatternSearch / xTZSearch (+ the refinement)
    def inter(s):
        s = once(s, '#include "InterSearch.h"\n', '#include "InterSearch.h"\n#include "CommonLib/cuda/VtmCudaME.h"   // libvtmme\n')
        # locals + one lambda, placed at the second "Do integer search" marker (the one inside xMotionEstimation)
        first = s.index("  //  Do integer search\n")
        second = s.index("  //  Do integer search\n", first + 1)
        s = s[:second] + """\
  // ---- libvtmme: the integer search (full search or TZ search) and the refinement that follows it run on the GPU
  vtmcuda::SearchOut cudaOut;
  bool cudaFracDone      = false;
  bool cudaIntRefineDone = false;
  // not on the GPU: weighted prediction, wrap-around / scaled references, and the two cases in which VTM rewrites a
  // reference's reconstruction plane in place after it was uploaded — sub-pictures treated as pictures
  // (extendSubPicBorder, EncSlice.cpp:1556; clipMv is then clipMvInSubpic) and the composite long-term reference
  // (EncGOP::updateCompositeReference; also the zeroMV / in-CTU branch of xPatternSearchFracDIF, InterSearch.cpp:4311)
  const bool cudaUsable  = vtmcuda::enabled() && !m_cDistParam.applyWeight && m_lumaClpRng.bd <= 10 && !wrap
                           && clipMv == clipMvInPic && !m_useCompositeRef && !cStruct.inCtuSearch
                           && !pu.cu->slice->getRefPic( eRefPicList, iRefIdxPred )->isRefScaled( pu.cs->pps )
                           && !pu.cs->sps->getWrapAroundEnabledFlag();
  auto cudaSearch = [&]( const bool tz, const bool tzFast, const Mv& tzStart )
  {
    vtmcuda::SearchIn in;
    memset( &in, 0, sizeof( in ) );   // identical inputs = identical bytes (the per-PU batching compares requests)
    in.refPic       = pu.cu->slice->getRefPic( eRefPicList, iRefIdxPred );
    in.x            = pu.Y().x;
    in.y            = pu.Y().y;
    in.w            = pu.Y().width;
    in.h            = pu.Y().height;
    in.org          = cStruct.pcPatternKey->buf;
    in.orgStride    = cStruct.pcPatternKey->stride;
    in.srLeft       = cStruct.searchRange.left;
    in.srRight      = cStruct.searchRange.right;
    in.srTop        = cStruct.searchRange.top;
    in.srBottom     = cStruct.searchRange.bottom;
    in.predQx       = predQuarter.getHor();
    in.predQy       = predQuarter.getVer();
    in.imvShift     = cStruct.imvShift;
    in.subShiftMode = cStruct.subShiftMode;
    in.bitDepth     = m_lumaClpRng.bd;
    in.useHad       = m_pcEncCfg->getUseHADME() && !pu.cs->slice->getDisableSATDForRD();
    in.useAltHpel   = cStruct.useAltHpelIf;
    in.lambdaMotion = m_pcRdCost->getSelectedMotionLambda();
    in.picW         = pu.cs->pps->getPicWidthInLumaSamples();
    in.picH         = pu.cs->pps->getPicHeightInLumaSamples();
    in.maxCuW       = pu.cs->sps->getMaxCUWidth();
    in.maxCuH       = pu.cs->sps->getMaxCUHeight();
    in.doFrac       = 0;
    if( !m_pcEncCfg->getMCTSEncConstraint() )
    {
      in.doFrac = ( pu.cu->imv == 0 || pu.cu->imv == IMV_HPEL ) ? 1 : 2;
    }
    if( in.doFrac == 2 )   // integer / 4-pel AMVR: xPatternSearchIntRefine follows the search on the GPU too
    {
      in.imv     = pu.cu->imv;
      in.numCand = amvpInfo.numCand;
      for( int i = 0; i < 2; i++ )
      {
        in.candX[i]      = amvpInfo.mvCand[i].getHor();
        in.candY[i]      = amvpInfo.mvCand[i].getVer();
        in.mvpIdxBits[i] = m_auiMVPIdxCost[i][AMVP_MAX_NUM_CANDS];
      }
      in.mvpIdx  = riMVPIdx;
      in.bits    = ruiBits;
      in.fWeight = fWeight;
    }
    in.tzSearch = tz;
    if( tz )   // xTZSearch's inputs (FastSearch=1/3; the re-search of a cached integer MV uses the fast settings)
    {
      in.tzExtended        = !tzFast && m_motionEstimationSearchMethod == MESEARCH_DIAMOND_ENHANCED;
      in.tzFast            = tzFast;
      in.tzSelective       = !tzFast && m_motionEstimationSearchMethod == MESEARCH_SELECTIVE;   // xTZSearchSelective (FastSearch=2)
      in.tzFirstSearchStop = m_pcEncCfg->getFastMEAssumingSmootherMVEnabled();
      in.tzStartX          = tzStart.getHor();
      in.tzStartY          = tzStart.getVer();
      in.tzSearchRange     = m_iSearchRange;
      in.tzNumSeeds        = m_uniMvListSize;
      for( int i = 0; i < m_uniMvListSize; i++ )
      {
        const BlkUniMvInfo* e = m_uniMvList + ( ( m_uniMvListIdx - 1 - i + m_uniMvListMaxSize ) % ( m_uniMvListMaxSize ) );
        in.tzSeedX[i] = e->uniMvs[eRefPicList][iRefIdxPred].getHor();
        in.tzSeedY[i] = e->uniMvs[eRefPicList][iRefIdxPred].getVer();
      }
    }
    vtmcuda::search( in, cudaOut );
    rcMv.set( cudaOut.mvX, cudaOut.mvY );
    ruiCost           = cudaOut.intSad;
    cudaFracDone      = in.doFrac == 1;
    cudaIntRefineDone = in.doFrac == 2;
  };
  // the TZ searches (diamond, enhanced diamond, selective) run on the GPU without hash ME / MCTS / composite reference
  const bool cudaTzUsable = cudaUsable && !m_pcEncCfg->getMCTSEncConstraint() && !m_pcEncCfg->getUseHashME()
                            && ( m_motionEstimationSearchMethod == MESEARCH_DIAMOND || m_motionEstimationSearchMethod == MESEARCH_DIAMOND_ENHANCED
                                 || ( m_motionEstimationSearchMethod == MESEARCH_SELECTIVE && m_iSearchRange <= 128 ) );
  if( vtmcuda::collecting() )   // collect pass: a search the GPU does not take is left to the real pass
  {
    const bool cudaFull = ( m_motionEstimationSearchMethod == MESEARCH_FULL ) || bBi || bQTBTMV;
    if( cudaFull ? !cudaUsable : !cudaTzUsable )
    {
      return;
    }
  }
""" + s[second:]
        s = once(s, "    xPatternSearch( cStruct, rcMv, ruiCost);\n", """\
    if( cudaUsable )
    {
      cudaSearch( false, false, Mv() );   // libvtmme: integer full search + refinement
      if( vtmcuda::collecting() ) return;   // collect pass of predInterSearch: the search runs in the PU's batch
    }
    else
    {
      xPatternSearch( cStruct, rcMv, ruiCost);
    }
""")
        s = once(s, "    xTZSearch(pu, eRefPicList, iRefIdxPred, cStruct, rcMv, ruiCost, NULL, false, true);\n", """\
    if( cudaTzUsable )
    {
      cudaSearch( true, true, rcMv );   // libvtmme: xTZSearch with the fast settings + refinement
      if( vtmcuda::collecting() ) return;
    }
    else
    {
      xTZSearch(pu, eRefPicList, iRefIdxPred, cStruct, rcMv, ruiCost, NULL, false, true);
    }
""")
        s = once(s, "    xPatternSearchFast(pu, eRefPicList, iRefIdxPred, cStruct, rcMv, ruiCost, pIntegerMv2Nx2NPred);\n", """\
    if( cudaTzUsable )
    {
      cudaSearch( true, false, rcMv );   // libvtmme: xTZSearch (FastSearch=1/3) + refinement
      if( vtmcuda::collecting() ) return;
    }
    else
    {
      xPatternSearchFast(pu, eRefPicList, iRefIdxPred, cStruct, rcMv, ruiCost, pIntegerMv2Nx2NPred);
    }
""")
        s = once(s, "    xPatternSearchFracDIF( pu, eRefPicList, iRefIdxPred, cStruct, rcMv, cMvHalf, cMvQter, ruiCost );\n", """\
    if( cudaFracDone )   // libvtmme: the GPU call above already refined this MV
    {
      cMvHalf.set( cudaOut.halfX, cudaOut.halfY );
      cMvQter.set( cudaOut.qterX, cudaOut.qterY );
      ruiCost = cudaOut.fracCost;
    }
    else
    {
      xPatternSearchFracDIF( pu, eRefPicList, iRefIdxPred, cStruct, rcMv, cMvHalf, cMvQter, ruiCost );
    }
""")
        s = once(s, "    xPatternSearchIntRefine( pu, cStruct, rcMv, rcMvPred, riMVPIdx, ruiBits, ruiCost, amvpInfo, fWeight);\n", """\
    if( cudaIntRefineDone )   // libvtmme: the GPU call above already ran xPatternSearchIntRefine
    {
      rcMv.set( cudaOut.amvrMvX, cudaOut.amvrMvY );
      riMVPIdx = cudaOut.mvpIdx;
      rcMvPred = amvpInfo.mvCand[riMVPIdx];
      m_pcRdCost->setCostScale( 0 );
      m_pcRdCost->setPredictor( rcMvPred );
      ruiBits  = cudaOut.bits;
      ruiCost  = cudaOut.cost;
    }
    else
    {
      xPatternSearchIntRefine( pu, cStruct, rcMv, rcMvPred, riMVPIdx, ruiBits, ruiCost, amvpInfo, fWeight);
    }
""")
        # 5. predInterSearch: before its (list, reference) loop of uni-predictive searches, the same loop in collect mode —
        # the searches of all reference pictures of the PU then run in ONE vtmme_search call (VtmCudaME.h)
        anchor = "      //  Uni-directional prediction\n      for ( int iRefList = 0; iRefList < iNumPredDir; iRefList++ )\n"
        pos = s.index(anchor, s.index("void InterSearch::predInterSearch("))
        s = s[:pos] + """\
      if( vtmcuda::enabled() && vtmcuda::batching() )   // libvtmme: collect the uni-predictive searches of this PU
      {
        vtmcuda::beginCollect();
        for( int cList = 0; cList < iNumPredDir; cList++ )
        {
          const RefPicList cRefPicList = cList ? REF_PIC_LIST_1 : REF_PIC_LIST_0;
          for( int cRefIdx = 0; cRefIdx < cs.slice->getNumRefIdx( cRefPicList ); cRefIdx++ )
          {
            if( m_pcEncCfg->getFastMEForGenBLowDelayEnabled() && cList == 1 && cs.slice->getList1IdxToList0Idx( cRefIdx ) >= 0 )
            {
              continue;   // the loop below copies list 0's result for this picture
            }
            uint32_t cBits = uiMbBits[cList];
            if( cs.slice->getNumRefIdx( cRefPicList ) > 1 )
            {
              cBits += cRefIdx + 1;
              if( cRefIdx == cs.slice->getNumRefIdx( cRefPicList ) - 1 )
              {
                cBits--;
              }
            }
            Mv         cPred, cMvOut;
            AMVPInfo   cAmvp;
            Distortion cDist = 0, cCost = 0;
            xEstimateMvPredAMVP( pu, origBuf, cRefPicList, cRefIdx, cPred, cAmvp, false, &cDist );
            int cMvpIdx = pu.mvpIdx[cRefPicList];
            cBits += m_auiMVPIdxCost[cMvpIdx][AMVP_MAX_NUM_CANDS];
            xMotionEstimation( pu, origBuf, cRefPicList, cPred, cRefIdx, cMvOut, cMvpIdx, cBits, cCost, cAmvp );
          }
        }
        vtmcuda::endCollect();
      }
""" + s[pos:]
        return s
    edit(os.path.join(lib, "EncoderLib", "InterSearch.cpp"), inter)

    # 6. EncTemporalFilter::filter: motion estimation of every neighbouring picture, applyMotion and the bilateral weighting on
    # the GPU (vtmcuda::temporalFilter); the weight tables are computed here, with the encoder's own exp()
    def tfilter(s):
        s = once(s, '#include "EncTemporalFilter.h"\n', '#include "EncTemporalFilter.h"\n#include "CommonLib/cuda/VtmCudaME.h"   // libvtmme\n#include <vector>\n')
        s = once(s, "      motionEstimation(srcPic.mvs, origPadded, srcPic.picBuffer, origSubsampled2, origSubsampled4);\n", """\
      if( !vtmcuda::enabled() )   // libvtmme: with the GPU path every neighbour is searched in one batch below
      {
        motionEstimation(srcPic.mvs, origPadded, srcPic.picBuffer, origSubsampled2, origSubsampled4);
      }
""")
        s = once(s, "    bilateralFilter(origPadded, srcFrameInfo, newOrgPic, overallStrength);\n", """\
    if( vtmcuda::enabled() && !srcFrameInfo.empty() && srcFrameInfo.size() <= 8 )   // libvtmme
    {
      const int numRefs = int( srcFrameInfo.size() ), numComp = getNumberValidComponents( m_chromaFormatIDC );
      int refStrengthRow = 2;
      if( numRefs == m_range * 2 )
      {
        refStrengthRow = 0;
      }
      else if( numRefs == m_range )
      {
        refStrengthRow = 1;
      }
      const double lumaSigmaSq = ( m_QP - m_sigmaZeroPoint ) * ( m_QP - m_sigmaZeroPoint ) * m_sigmaMultiplier;
      const double chromaSigmaSq = 30 * 30;
      std::vector<vtmcuda::TfPlane> org( numComp ), refs( numRefs * numComp );
      std::vector<std::vector<double>> tables( numComp );
      std::vector<const double*> weights( numComp );
      std::vector<int16_t*> dst( numComp );
      std::vector<int> dstStride( numComp ), bitDepth( numComp );
      for( int c = 0; c < numComp; c++ )
      {
        const ComponentID compID = ComponentID( c );
        org[c] = vtmcuda::TfPlane{ origPadded.bufs[c].buf, int( origPadded.bufs[c].stride ), int( origPadded.bufs[c].width ), int( origPadded.bufs[c].height ) };
        for( int i = 0; i < numRefs; i++ )
        {
          const PelBuf& rb = srcFrameInfo[i].picBuffer.bufs[c];
          refs[i * numComp + c] = vtmcuda::TfPlane{ rb.buf, int( rb.stride ), int( rb.width ), int( rb.height ) };
        }
        dst[c]       = newOrgPic.bufs[c].buf;
        dstStride[c] = int( newOrgPic.bufs[c].stride );
        bitDepth[c]  = m_internalBitDepth[toChannelType( compID )];
        // the weight of a neighbouring sample as a function of |refVal - orgVal| — the expressions of bilateralFilter
        const double sigmaSq = isChroma( compID ) ? chromaSigmaSq : lumaSigmaSq;
        const double weightScaling = overallStrength * ( isChroma( compID ) ? m_chromaFactor : 0.4 );
        const Pel maxSampleValue = ( 1 << bitDepth[c] ) - 1;
        const double bitDepthDiffWeighting = 1024.0 / ( maxSampleValue + 1 );
        tables[c].resize( size_t( numRefs ) << bitDepth[c] );
        for( int i = 0; i < numRefs; i++ )
        {
          const int index = std::min( 1, std::abs( srcFrameInfo[i].origOffset ) - 1 );
          for( int d = 0; d <= maxSampleValue; d++ )
          {
            double diff = (double) d;
            diff *= bitDepthDiffWeighting;
            double diffSq = diff * diff;
            tables[c][( size_t( i ) << bitDepth[c] ) + d] = weightScaling * m_refStrengths[refStrengthRow][index] * exp( -diffSq / ( 2 * sigmaSq ) );
          }
        }
        weights[c] = tables[c].data();
      }
      vtmcuda::temporalFilter( numRefs, numComp, org.data(), refs.data(), getComponentScaleX( COMPONENT_Cb, m_chromaFormatIDC ),
                               getComponentScaleY( COMPONENT_Cb, m_chromaFormatIDC ), bitDepth.data(), weights.data(), dst.data(), dstStride.data() );
    }
    else
    {
      if( vtmcuda::enabled() )   // not taken by the GPU path: the motion estimation skipped above
      {
        for( auto& srcPic : srcFrameInfo )
        {
          motionEstimation( srcPic.mvs, origPadded, srcPic.picBuffer, origSubsampled2, origSubsampled4 );
        }
      }
      bilateralFilter(origPadded, srcFrameInfo, newOrgPic, overallStrength);
    }
""")
        return s
    edit(os.path.join(lib, "EncoderLib", "EncTemporalFilter.cpp"), tfilter)
    print("patched tree:", dst)


if __name__ == "__main__":
    main()
