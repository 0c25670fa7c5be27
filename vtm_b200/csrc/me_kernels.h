// Host-side launch interface of the motion-search kernels (internal to libvtmme.so).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "me_common.cuh"

namespace vtmme {

// Dynamic shared memory above 48 KB is an opt-in per kernel AND per device: launchers remember what they configured for
// the current device (a process may hold contexts on several GPUs).
struct SmemOptIn
{
  size_t configured[64] = {};
  template <class Kernel>
  cudaError_t ensure(Kernel k, size_t bytes, size_t floorBytes = 0)
  {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    size_t& c = configured[dev & 63];
    if (c < floorBytes) c = floorBytes;
    if (bytes <= c) return cudaSuccess;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
    if (e == cudaSuccess) c = bytes;
    return e;
  }
};

// Geometry of the batched per-CTU search (vtmme_search_frames): 5 levels of grid-aligned square CUs.
struct FrameGeom
{
  int picW, picH;
  int nx[5], ny[5];   // CUs per level (size 8<<l), floor(pic/size)
  int off[6];         // level offsets in the CU order, off[5] = total
  int nRegX, nRegY;   // 32x32 regions, ceil
  int nCtuX, nCtuY;   // 128x128 CTUs, ceil
};

inline FrameGeom make_geom(int w, int h)
{
  FrameGeom g;
  g.picW = w;
  g.picH = h;
  int acc = 0;
  for (int l = 0; l < 5; l++)
  {
    g.nx[l]  = w / (8 << l);
    g.ny[l]  = h / (8 << l);
    g.off[l] = acc;
    acc += g.nx[l] * g.ny[l];
  }
  g.off[5] = acc;
  g.nRegX  = (w + 31) / 32;
  g.nRegY  = (h + 31) / 32;
  g.nCtuX  = (w + 127) / 128;
  g.nCtuY  = (h + 127) / 128;
  return g;
}

struct TreeParams
{
  FrameGeom           g;
  const DevPic*       cur;       // [nPairs]
  const DevPic*       ref;       // [nPairs]
  const short2*       predQ;     // [nPairs][nCU] quarter-pel, or nullptr (zero)
  unsigned long long* keys;      // [nPairs][nCU] best (cost, position) per CU
  uint32_t*           surf;      // [nPairs][nReg][surfCap] SAD32 surfaces for the 64/128 levels
  uint32_t*           surfEven;  // same, even rows only x2 (subShiftMode 2: feeds the 64x64 level); nullptr in mode 0
  int4*               regInfo;   // [nPairs][nReg] {wl8, wt, ngx, nrows} of each surface
  int*                errFlag;
  int                 surfCap;     // elements per region surface
  int                 maxGx;       // capacity: 8-wide displacement groups per row
  int                 maxRows;     // capacity: displacement rows per region
  int                 bandRows;    // displacement rows staged per band
  int                 sr, ctu, imvShift;
  int                 subShiftMode;   // 0 or 2
  double              lambda;
};

size_t tree_sad_smem_bytes(int maxGx, int maxRows, int bandRows);
int    tree_band_override();   // development knob (VTMME_TREE_VARIANT), 0 = none
int    tree_pick_band_rows(int maxGx, int maxRows, bool subSampling);
cudaError_t launch_tree_sad(const TreeParams& p, int nPairs, cudaStream_t st);
cudaError_t launch_tree_upper(const TreeParams& p, int nPairs, cudaStream_t st);

// TZ search of the frame path (me_tz.cu): fills keys like the tree kernels do
struct TzFrameParams
{
  FrameGeom           g;
  const DevPic*       cur;
  const DevPic*       ref;
  const short2*       predQ;
  unsigned long long* keys;
  int                 sr, ctu, imvShift, subShiftMode;
  int                 extended, firstSearchStop;
  int                 selective;   // FastSearch=2: xTZSearchSelective; subShiftMode 1 = staged SAD
  double              lambda;
};
cudaError_t launch_tz_frame(const TzFrameParams& p, int nPairs, int predSpread, cudaStream_t st, int* launches);
// me_tz_smem.cu: one level (0 .. 3) with the search windows staged in shared memory; *done = false when they do not fit
cudaError_t launch_tz_frame_smem(const TzFrameParams& p, int level, int nPairs, int predSpread, cudaStream_t st, bool* done);

// Fractional refinement + result write-out for the frame path.
struct FracFrameParams
{
  FrameGeom                 g;
  const DevPic*             cur;
  const DevPic*             ref;
  const short2*             predQ;
  const unsigned long long* keys;
  void*                     results;   // vtmme_cu_result[nPairs][nCU]
  int                       bitDepth, imvShift, useHad, fracMode;
  double                    lambda;
};
size_t      frac_frame_acc_bytes(const FrameGeom& g, int nPairs);
cudaError_t launch_frac_frame(const FracFrameParams& p, uint32_t* acc, int nPairs, cudaStream_t st, int* launches);
// me_frac_tile.cu: one thread per 8x8 SATD tile, one launch per CU level (what launch_frac_frame uses when fracMode != 0)
cudaError_t launch_frac_frame_tiles(const FracFrameParams& p, int nPairs, cudaStream_t st, int* launches);

// State of an xPatternSearchIntRefine call (InterSearch.cpp:4172-4282), MVs in 1/16 sample
struct DevAmvr
{
  int      imv, numCand;
  int      candX[2], candY[2];
  int      mvpIdx;
  uint32_t mvpIdxBits[2];
  uint32_t bits;
  int      posX, posY, picW, picH, maxCuW, maxCuH;   // clipMv
  double   fWeight;
};

// State of an xTZSearch call (InterSearch.cpp:3640-3974), see vtmme_tz in include/vtmme.h
struct DevTz
{
  int startX, startY;
  int hasInt2Nx2N, int2Nx2NX, int2Nx2NY;
  int nSeeds;
  int seedX[16], seedY[16];
  int searchRange;
  int extended, fast, firstSearchStop;
  int posX, posY, picW, picH, maxCuW, maxCuH;
  int selective;   // xTZSearchSelective instead of xTZSearch
  int staged;      // cStruct.subShiftMode == 1: the probes use xTZSearchHelp's staged SAD
};

// Generic per-call jobs (vtmme_search)
struct DevJob
{
  const int16_t* org;        // device pointer to the pattern
  int            orgStride;
  const int16_t* refAtPU;    // device pointer: reference plane at the PU position
  int            refStride;
  int            w, h;
  int            l, r, t, b;
  int            predQx, predQy;
  int            imvShift, subShift, bitDepth, useHad, useAltHpel, fracMode;
  int            signedOrg;  // pattern may be outside [0, 2^bd): bi-pred 2*org - otherPred
  double         lambda;
  DevAmvr        amvr;       // fracMode 2 only
};
struct DevJobResult
{
  int                mvX, mvY;
  unsigned long long intSad;
  int                halfX, halfY, qterX, qterY;
  unsigned long long fracCost;
  int                amvrMvX, amvrMvY, mvpIdx;   // fracMode 2: xPatternSearchIntRefine's rcMv (1/16), riMVPIdx,
  uint32_t           bits;                       //   ruiBits and
  unsigned long long cost;                       //   ruiCost
};
cudaError_t launch_job_search_impl(const DevJob* dJobs, unsigned long long* dKeys, DevJobResult* dResults, int n,
                                   int maxRegions, int nSplit, int bandRows, int maxGx, bool anyMulti, uint32_t* dSurf,
                                   const long long* dSurfOff, uint32_t* dFracAcc, int maxFracChunks, cudaStream_t st,
                                   int* launches, const DevTz* dTz = nullptr, int maxPatternSamples = 0,
                                   unsigned int* done = nullptr, unsigned int seq = 0, bool* fusedTz = nullptr);

// One small job per launch (me_job_fused_kernel): descriptor and pattern as kernel parameters
struct FusedJobArgs
{
  DevJob              job;             // job.org is used only when !inlinePattern (pattern read from an uploaded picture)
  unsigned long long* key;             // persistent slot, all-ones between calls
  DevJobResult*       result;          // mapped pinned host memory
  unsigned int*       ticket;          // zero between calls
  unsigned int*       done;            // mapped pinned host memory: set to `seq` after the result is visible (or nullptr)
  unsigned int        seq;
  int                 bandRows;        // window rows per CTA pass
  int                 inlinePattern;
  int16_t             pattern[32 * 32];   // row stride job.w
};
size_t      fused_job_smem_bytes(const FusedJobArgs& a);
cudaError_t launch_job_fused(const FusedJobArgs& a, int grid, cudaStream_t st);

// Table-level batches
cudaError_t launch_dist_batch(int kind, const int16_t* org, int orgStride, long long orgBlockStride, const int16_t* cur,
                              int curStride, long long curBlockStride, int w, int h, int subShift, int n,
                              unsigned long long* out, cudaStream_t st);
cudaError_t launch_interp_batch(int comp, int vertical, const int16_t* src, int srcStride, long long srcBlockStride,
                                int16_t* dst, int dstStride, long long dstBlockStride, int w, int h, int frac, int isFirst,
                                int isLast, int bitDepth, int useAltHpel, int n, cudaStream_t st);
cudaError_t launch_filter_batch(int taps, int vertical, int isFirst, int isLast, int copy, const int16_t* src, int srcStride,
                                long long srcBlockStride, int16_t* dst, int dstStride, long long dstBlockStride, int w, int h,
                                const int16_t* coeff, int bitDepth, int n, cudaStream_t st);
cudaError_t launch_extend_border(DevPic pic, cudaStream_t st);
cudaError_t launch_scatter_extend(DevPic pic, const int16_t* staging, int srcStride, cudaStream_t st);

// Motion compensation (mc_kernels.cu): a block is cut into tiles of at most 16x16 outputs, one warp each
struct McTile
{
  const int16_t* src;          // reference plane at the tile's integer position (block position + (mv >> shift))
  int16_t*       dst;
  int            srcStride, dstStride;
  uint8_t        tw, th;       // 1..16
  uint8_t        xFrac, yFrac; // 1/16 (luma) or 1/32 (chroma) phase
  uint8_t        q4Hor, q4Ver; // the 4x4 coefficient table applies to the horizontal / vertical pass
  // DMVR's padded prediction (xFinalPaddedMCForDMVR): winMaxX != 0 makes src the origin of the prefetched window, the tile
  // starts at window coordinates (winX, winY) and every read is clamped to [0, winMaxX] x [0, winMaxY] (xPad's replication)
  uint8_t        winMaxX, winMaxY;
  int8_t         winX, winY;
  uint8_t        pad2[4];
};
// the same tile with the original block it is compared against (candidate SAD, mc_sad_kernel)
struct McSadTile
{
  McTile         mc;          // mc.dst unused
  const int16_t* org;         // original block at the tile position
  int            orgStride;
  int            outIdx;      // candidate slot the tile's partial SAD is added to
  int            subShift;    // rows with (row & ((1 << subShift) - 1)) == 0 are summed, the sum is shifted back up
  int            pad;
};
cudaError_t launch_mc_sad(const McSadTile* dTiles, int nTiles, int bitDepth, int useAltHpel, unsigned long long* dOut,
                          cudaStream_t st);
cudaError_t launch_mc_batch(int comp, const McTile* dTiles, int nTiles, int bi, int bitDepth, int useAltHpel, cudaStream_t st);
cudaError_t launch_add_avg(const int16_t* a, const int16_t* b, int16_t* d, long long n, int bitDepth, cudaStream_t st);
cudaError_t launch_remove_high_freq(int16_t* d, const int16_t* s, long long n, int clip, int bitDepth, cudaStream_t st);
cudaError_t launch_add_weighted_avg(const int16_t* a, const int16_t* b, int16_t* d, long long n, int bitDepth, int bcwIdx, cudaStream_t st);
cudaError_t launch_remove_weight_high_freq(int16_t* d, const int16_t* s, long long n, int clip, int bitDepth, int bcwWeight, cudaStream_t st);

// Symmetric-MVD search (smvd_kernels.cu): state of one InterSearch::xSymmetricMotionEstimation call, MVs in 1/16 sample
struct DevSmvd
{
  const int16_t*     org;          // original block at the PU position (device)
  const int16_t*     refCur;       // sample (0,0) of the searched list's reference plane
  const int16_t*     refTar;       // ... of the other list's
  int                orgStride, refStride;
  int                x, y, w, h, picW, picH, maxCu, bd, imv;
  int                curPredX, curPredY, tarPredX, tarPredY;
  int                curMvX, curMvY, tarMvX, tarMvY;
  int                clipBiPred, useHad, bcwIdx;
  double             lambda;
  unsigned long long cost;
};
struct DevSmvdResult
{
  int                curMvX, curMvY, tarMvX, tarMvY;
  unsigned long long cost;
};
cudaError_t launch_smvd_search(const DevSmvd* dJobs, DevSmvdResult* dOut, int n, cudaStream_t st);

// Affine ME primitives (affine_kernels.cu)
struct DevAffineBlock
{
  const int16_t* org;
  const int16_t* pred;
  int            orgStride, predStride, w, h, sixParam, pad;
};
cudaError_t launch_affine_sobel(const int16_t* pred, int stride, int w, int h, int vertical, int* deriv, cudaStream_t st);
cudaError_t launch_affine_equal_coeff(const int16_t* residue, int residueStride, const int* d0, const int* d1, int derivStride, int w, int h,
                                      int sixParam, long long* coeff, cudaStream_t st);
cudaError_t launch_affine_step(const DevAffineBlock* dBlocks, int n, long long* dCoeff, cudaStream_t st);

// GOP-based temporal filter motion estimation (mctf_kernels.cu)
struct MctfLevelParams
{
  const DevPic* org;        // [nPairs] original picture of this level
  const DevPic* ref;        // [nPairs] reference ("buffer") picture of this level
  int           width, height;   // picture size of this level
  const int3*   previous;   // [nPairs][prevH][prevW] vectors of the coarser level, or nullptr
  int           prevW, prevH, factor;
  int3*         mvs;        // [nPairs][mvH][mvW] {x, y, error}
  int           mvW, mvH;
  int           maxv;       // (1 << bitDepth) - 1
};
cudaError_t launch_mctf_subsample(DevPic in, DevPic out, cudaStream_t st);
cudaError_t launch_mctf_init_mv(int3* mv, int n, cudaStream_t st);
cudaError_t launch_mctf_apply_motion(DevPic src, int csx, int csy, const int3* mv, int mvStride, int maxv, int16_t* dst,
                                     cudaStream_t st);
cudaError_t launch_mctf_level(const MctfLevelParams& p, int blockSize, bool doubleRes, int nPairs, cudaStream_t st);
cudaError_t launch_mctf_bilateral(DevPic org, const DevPic* corr, int numRefs, const double* dWeights, int bitDepth, int16_t* dst,
                                  cudaStream_t st);

// Decoder-side MV refinement of a batch of sub-blocks (dmvr_kernels.cu); same layouts as vtmme_dmvr_block / _result
struct DevDmvrBlock
{
  int x, y, w, h;
  int mv0x, mv0y, mv1x, mv1y;
};
struct DevDmvrResult
{
  int      mvdX, mvdY;
  uint32_t minCost;
  int      notZeroCost;
};
cudaError_t launch_dmvr_refine(const DevPic& ref0, const DevPic& ref1, const DevDmvrBlock* blocks, int n, int bitDepth, int maxCu,
                               DevDmvrResult* results, cudaStream_t st);

}   // namespace vtmme
