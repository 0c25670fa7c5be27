/* TEST INFRASTRUCTURE — CPU oracle, not part of the product (see vtm_oracle.h for the rules and the
 * parity status: PINNED against the compiled reference, oracle/_ref/libvtmref.so).
 *
 * Every function cites the VTM 9.3 file:line it restates (paths relative to /root/reference/source/Lib).
 */
#include "vtm_oracle.h"

#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define VO_MAX_CU 128
#define VO_NTAPS_LUMA 8
#define VO_IF_INTERNAL_PREC 14 /* InterpolationFilter.h:48 */
#define VO_IF_FILTER_PREC 6    /* :49 */
#define VO_IF_INTERNAL_OFFS (1 << (VO_IF_INTERNAL_PREC - 1))

static int vo_floor_log2(uint32_t x) /* CommonDef.h:640-648 */
{
  int r = -1;
  while (x)
  {
    x >>= 1;
    r++;
  }
  return r;
}

/* ------------------------------------------------------------------------------------------------
 * Distortion
 * ---------------------------------------------------------------------------------------------- */

/* RdCost::setDistParam, sub-sampling choice — CommonLib/RdCost.cpp:289-323 */
int vo_subshift(int subShiftMode, int w, int h)
{
  int s = 0;
  if (subShiftMode == 1)
  {
    if (h > 32 && (h & 15) == 0)
      s = 4;
    else if (h > 16 && (h & 7) == 0)
      s = 3;
    else if (h > 8 && (h & 3) == 0)
      s = 2;
    else if ((h & 1) == 0)
      s = 1;
  }
  else if (subShiftMode == 2)
  {
    if (h > 8 && w <= 64) s = 1;
  }
  else if (subShiftMode == 3)
  {
    if (h > 8) s = 1;
  }
  return s;
}

/* RdCost::xGetSAD — CommonLib/RdCost.cpp:493-528 (the SIMD versions x86/RdCostX86.h:210-456 compute the
 * same sum; the scalar early exit on maximumDistortionForEarlyExit :518-521 never changes an argmin and
 * is not restated).  DISTORTION_PRECISION_ADJUSTMENT is 0 (TypeDef.h:228-233). */
uint64_t vo_sad(const vo_pel* org, int orgStride, const vo_pel* cur, int curStride, int w, int h, int subShift)
{
  const int step = 1 << subShift;
  uint64_t  sum  = 0;
  for (int rows = h; rows != 0; rows -= step)
  {
    for (int n = 0; n < w; n++) sum += (uint64_t) abs(org[n] - cur[n]);
    org += orgStride * step;
    cur += curStride * step;
  }
  return sum << subShift;
}

/* One Hadamard tile: Σ|coef| with the DC term scaled >>2 (JVET_R0164_MEAN_SCALED_SATD, TypeDef.h:62).
 * Restates the butterflies of xCalcHADs{2x2,4x4,8x8,16x8,8x16,4x8,8x4} — CommonLib/RdCost.cpp:2140-2817.
 * The reference's fixed butterfly networks are Walsh-Hadamard transforms along rows and columns; the
 * sum of absolute coefficients does not depend on the output ordering, only the DC position matters
 * (index [0][0] in every reference variant). */
static int64_t vo_had_tile_raw(const vo_pel* org, int orgStride, const vo_pel* cur, int curStride, int tw, int th)
{
  int m[16][16];
  for (int y = 0; y < th; y++)
    for (int x = 0; x < tw; x++) m[y][x] = org[y * orgStride + x] - cur[y * curStride + x];
  for (int y = 0; y < th; y++) /* horizontal */
    for (int len = 1; len < tw; len <<= 1)
      for (int i = 0; i < tw; i += len << 1)
        for (int k = i; k < i + len; k++)
        {
          int a = m[y][k], b = m[y][k + len];
          m[y][k]       = a + b;
          m[y][k + len] = a - b;
        }
  for (int x = 0; x < tw; x++) /* vertical */
    for (int len = 1; len < th; len <<= 1)
      for (int i = 0; i < th; i += len << 1)
        for (int k = i; k < i + len; k++)
        {
          int a = m[k][x], b = m[k + len][x];
          m[k][x]       = a + b;
          m[k + len][x] = a - b;
        }
  int64_t s = 0;
  for (int y = 0; y < th; y++)
    for (int x = 0; x < tw; x++) s += abs(m[y][x]);
  s -= abs(m[0][0]);
  s += abs(m[0][0]) >> 2;
  return s;
}

static uint64_t vo_had_tile(const vo_pel* org, int os, const vo_pel* cur, int cs, int tw, int th)
{
  int64_t s = vo_had_tile_raw(org, os, cur, cs, tw, th);
  if (tw == 2 && th == 2) return (uint64_t) s;                 /* RdCost.cpp:2154-2163 */
  if (tw == 4 && th == 4) return (uint64_t) ((s + 1) >> 1);    /* :2262 */
  if (tw == 8 && th == 8) return (uint64_t) ((s + 2) >> 2);    /* :2363 */
  if (tw * th == 128) return (uint64_t) (int) ((int) s / sqrt(16.0 * 8) * 2); /* :2513, :2654 */
  return (uint64_t) (int) ((int) s / sqrt(4.0 * 8) * 2);       /* :2731, :2814 */
}

/* RdCost::xGetHADs — CommonLib/RdCost.cpp:2819-2934 (same tiling as xGetHADs_SIMD,
 * x86/RdCostX86.h:2154-2291; its AVX2 16x16 tile is four 8x8 tiles). */
uint64_t vo_satd(const vo_pel* org, int os, const vo_pel* cur, int cs, int w, int h)
{
  int tw, th;
  if (w > h && (h & 7) == 0 && (w & 15) == 0)
    tw = 16, th = 8;
  else if (w < h && (w & 7) == 0 && (h & 15) == 0)
    tw = 8, th = 16;
  else if (w > h && (h & 3) == 0 && (w & 7) == 0)
    tw = 8, th = 4;
  else if (w < h && (w & 3) == 0 && (h & 7) == 0)
    tw = 4, th = 8;
  else if ((h % 8 == 0) && (w % 8 == 0))
    tw = 8, th = 8;
  else if ((h % 4 == 0) && (w % 4 == 0))
    tw = 4, th = 4;
  else
    tw = 2, th = 2;
  uint64_t sum = 0;
  for (int y = 0; y < h; y += th)
    for (int x = 0; x < w; x += tw) sum += vo_had_tile(org + y * os + x, os, cur + y * cs + x, cs, tw, th);
  return sum;
}

/* ------------------------------------------------------------------------------------------------
 * Motion-vector rate
 * ---------------------------------------------------------------------------------------------- */

/* RdCost::xGetExpGolombNumberOfBits — CommonLib/RdCost.h:301-313 (MAX_CU_SIZE 128, MAX_CU_DEPTH 7) */
uint32_t vo_eg_bits(int v)
{
  unsigned len = 1, t = (v <= 0) ? ((unsigned) (-v) << 1) + 1 : (unsigned) (v << 1);
  while (t > 128)
  {
    len += 14;
    t >>= 7;
  }
  return len + ((unsigned) vo_floor_log2(t) << 1);
}

/* RdCost::getBitsOfVectorWithPredictor — CommonLib/RdCost.h:315 */
uint32_t vo_mv_bits(int x, int y, int predX, int predY, int costScale, int imvShift)
{
  return vo_eg_bits(((x << costScale) - predX) >> imvShift) + vo_eg_bits(((y << costScale) - predY) >> imvShift);
}

/* RdCost::getCost — CommonLib/RdCost.h:191 */
uint64_t vo_mv_cost(double lambdaMotion, uint32_t bits) { return (uint64_t) (lambdaMotion * bits); }

/* ------------------------------------------------------------------------------------------------
 * Interpolation
 * ---------------------------------------------------------------------------------------------- */

/* CommonLib/InterpolationFilter.cpp:57-95, 181-216 */
static const int16_t vo_luma4x4[16][8] = {
  { 0, 0, 0, 64, 0, 0, 0, 0 },     { 0, 1, -3, 63, 4, -2, 1, 0 },   { 0, 1, -5, 62, 8, -3, 1, 0 },
  { 0, 2, -8, 60, 13, -4, 1, 0 },  { 0, 3, -10, 58, 17, -5, 1, 0 }, { 0, 3, -11, 52, 26, -8, 2, 0 },
  { 0, 2, -9, 47, 31, -10, 3, 0 }, { 0, 3, -11, 45, 34, -10, 3, 0 }, { 0, 3, -11, 40, 40, -11, 3, 0 },
  { 0, 3, -10, 34, 45, -11, 3, 0 }, { 0, 3, -10, 31, 47, -9, 2, 0 }, { 0, 2, -8, 26, 52, -11, 3, 0 },
  { 0, 1, -5, 17, 58, -10, 3, 0 }, { 0, 1, -4, 13, 60, -8, 2, 0 },  { 0, 1, -3, 8, 62, -5, 1, 0 },
  { 0, 1, -2, 4, 63, -3, 1, 0 }
};
static const int16_t vo_luma[16][8] = {
  { 0, 0, 0, 64, 0, 0, 0, 0 },       { 0, 1, -3, 63, 4, -2, 1, 0 },     { -1, 2, -5, 62, 8, -3, 1, 0 },
  { -1, 3, -8, 60, 13, -4, 1, 0 },   { -1, 4, -10, 58, 17, -5, 1, 0 },  { -1, 4, -11, 52, 26, -8, 3, -1 },
  { -1, 3, -9, 47, 31, -10, 4, -1 }, { -1, 4, -11, 45, 34, -10, 4, -1 }, { -1, 4, -11, 40, 40, -11, 4, -1 },
  { -1, 4, -10, 34, 45, -11, 4, -1 }, { -1, 4, -10, 31, 47, -9, 3, -1 }, { -1, 3, -8, 26, 52, -11, 4, -1 },
  { 0, 1, -5, 17, 58, -10, 4, -1 },  { 0, 1, -4, 13, 60, -8, 3, -1 },   { 0, 1, -3, 8, 62, -5, 2, -1 },
  { 0, 1, -2, 4, 63, -3, 1, 0 }
};
static const int16_t vo_luma_alt_hpel[8] = { 0, 3, 9, 20, 20, 9, 3, 0 };
static const int16_t vo_chroma[32][4]    = {
  { 0, 64, 0, 0 },    { -1, 63, 2, 0 },   { -2, 62, 4, 0 },   { -2, 60, 7, -1 },  { -2, 58, 10, -2 }, { -3, 57, 12, -2 },
  { -4, 56, 14, -2 }, { -4, 55, 15, -2 }, { -4, 54, 16, -2 }, { -5, 53, 18, -2 }, { -6, 52, 20, -2 }, { -6, 49, 24, -3 },
  { -6, 46, 28, -4 }, { -5, 44, 29, -4 }, { -4, 42, 30, -4 }, { -4, 39, 33, -4 }, { -4, 36, 36, -4 }, { -4, 33, 39, -4 },
  { -4, 30, 42, -4 }, { -4, 29, 44, -5 }, { -4, 28, 46, -6 }, { -3, 24, 49, -6 }, { -2, 20, 52, -6 }, { -2, 18, 53, -5 },
  { -2, 16, 54, -4 }, { -2, 15, 55, -4 }, { -2, 14, 56, -4 }, { -2, 12, 57, -3 }, { -2, 10, 58, -2 }, { -1, 7, 60, -2 },
  { 0, 4, 62, -2 },   { 0, 2, 63, -1 }
};

static int vo_clip(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* InterpolationFilter::filterCopy<isFirst,isLast> — CommonLib/InterpolationFilter.cpp:397-524 (biMCForDMVR=false) */
static void vo_filter_copy(int isFirst, int isLast, int bd, const vo_pel* src, int ss, vo_pel* dst, int ds, int w, int h)
{
  const int shift = (VO_IF_INTERNAL_PREC - bd) > 2 ? (VO_IF_INTERNAL_PREC - bd) : 2;
  const int maxv  = (1 << bd) - 1;
  for (int row = 0; row < h; row++)
  {
    for (int col = 0; col < w; col++)
    {
      if (isFirst == isLast)
        dst[col] = src[col];
      else if (isFirst)
      {
        vo_pel val = (vo_pel) (src[col] << shift);
        dst[col]   = (vo_pel) (val - (vo_pel) VO_IF_INTERNAL_OFFS);
      }
      else
      {
        vo_pel val = src[col];
        val        = (vo_pel) ((val + VO_IF_INTERNAL_OFFS + (1 << (shift - 1))) >> shift);
        dst[col]   = (vo_pel) vo_clip(val, 0, maxv);
      }
    }
    src += ss;
    dst += ds;
  }
}

/* InterpolationFilter::filter<N,isVertical,isFirst,isLast> — CommonLib/InterpolationFilter.cpp:550-656 */
static void vo_filter(int N, int isVertical, int isFirst, int isLast, int bd, const vo_pel* src, int ss, vo_pel* dst,
                      int ds, int w, int h, const int16_t* coeff)
{
  const int cStride = isVertical ? ss : 1;
  src -= (N / 2 - 1) * cStride;
  const int headRoom = (VO_IF_INTERNAL_PREC - bd) > 2 ? (VO_IF_INTERNAL_PREC - bd) : 2;
  int       shift    = VO_IF_FILTER_PREC;
  int       offset;
  const int maxv = (1 << bd) - 1;
  if (isLast)
  {
    shift += isFirst ? 0 : headRoom;
    offset = 1 << (shift - 1);
    offset += isFirst ? 0 : VO_IF_INTERNAL_OFFS << VO_IF_FILTER_PREC;
  }
  else
  {
    shift -= isFirst ? headRoom : 0;
    offset = isFirst ? -(VO_IF_INTERNAL_OFFS << shift) : 0;
  }
  for (int row = 0; row < h; row++)
  {
    for (int col = 0; col < w; col++)
    {
      int sum = 0;
      for (int k = 0; k < N; k++) sum += src[col + k * cStride] * coeff[k];
      vo_pel val = (vo_pel) ((sum + offset) >> shift);
      if (isLast) val = (vo_pel) vo_clip(val, 0, maxv);
      dst[col] = val;
    }
    src += ss;
    dst += ds;
  }
}

/* InterpolationFilter::filterHor (public dispatch) — CommonLib/InterpolationFilter.cpp:749-810, nFilterIdx 0 */
void vo_filter_hor(int comp, const vo_pel* src, int ss, vo_pel* dst, int ds, int w, int h, int frac, int isLast, int bd,
                   int useAltHpel)
{
  if (frac == 0)
    vo_filter_copy(1, isLast, bd, src, ss, dst, ds, w, h);
  else if (comp == 0)
  {
    const int16_t* c;
    if (frac == 8 && useAltHpel)
      c = vo_luma_alt_hpel;
    else if ((w == 4 && h == 4) || (w == 4 && h == (4 + VO_NTAPS_LUMA - 1)))
      c = vo_luma4x4[frac];
    else
      c = vo_luma[frac];
    vo_filter(8, 0, 1, isLast, bd, src, ss, dst, ds, w, h, c);
  }
  else
    vo_filter(4, 0, 1, isLast, bd, src, ss, dst, ds, w, h, vo_chroma[frac]); /* 4:2:0: frac << (1-csx), csx = 1 */
}

/* InterpolationFilter::filterVer (public dispatch) — CommonLib/InterpolationFilter.cpp:829-895 */
void vo_filter_ver(int comp, const vo_pel* src, int ss, vo_pel* dst, int ds, int w, int h, int frac, int isFirst,
                   int isLast, int bd, int useAltHpel)
{
  if (frac == 0)
    vo_filter_copy(isFirst, isLast, bd, src, ss, dst, ds, w, h);
  else if (comp == 0)
  {
    const int16_t* c;
    if (frac == 8 && useAltHpel)
      c = vo_luma_alt_hpel;
    else if (w == 4 && h == 4)
      c = vo_luma4x4[frac];
    else
      c = vo_luma[frac];
    vo_filter(8, 1, isFirst, isLast, bd, src, ss, dst, ds, w, h, c);
  }
  else
    vo_filter(4, 1, isFirst, isLast, bd, src, ss, dst, ds, w, h, vo_chroma[frac]);
}

/* ------------------------------------------------------------------------------------------------
 * Search window
 * ---------------------------------------------------------------------------------------------- */

/* clipMvInPic — CommonLib/Mv.cpp:53-71; InterSearch::xClipMv — EncoderLib/InterSearch.cpp:7735-7764
 * (no wrap-around, no sub-pictures: identical bounds) */
void vo_clip_mv(int* mvx, int* mvy, int posX, int posY, int picW, int picH, int maxCuW, int maxCuH)
{
  const int mvShift = 4, offset = 8;
  int       horMax = (picW + offset - posX - 1) << mvShift;
  int       horMin = (-maxCuW - offset - posX + 1) * (1 << mvShift);
  int       verMax = (picH + offset - posY - 1) << mvShift;
  int       verMin = (-maxCuH - offset - posY + 1) * (1 << mvShift);
  *mvx             = *mvx > horMax ? horMax : (*mvx < horMin ? horMin : *mvx);
  *mvy             = *mvy > verMax ? verMax : (*mvy < verMin ? verMin : *mvy);
}

static int vo_div_pow2(int v, int i) /* Mv::divideByPowerOf2 — CommonLib/Mv.h:128-136 */
{
  const int offset = 1 << (i - 1);
  return (v + offset - (v >= 0)) >> i;
}

/* InterSearch::xSetSearchRange — EncoderLib/InterSearch.cpp:3496-3535 */
void vo_set_search_range(int predX16, int predY16, int posX, int posY, int picW, int picH, int maxCuW, int maxCuH,
                         int searchRange, int* left, int* right, int* top, int* bottom)
{
  int px = predX16, py = predY16;
  vo_clip_mv(&px, &py, posX, posY, picW, picH, maxCuW, maxCuH);
  int tlx = px - (searchRange << 4), tly = py - (searchRange << 4);
  int brx = px + (searchRange << 4), bry = py + (searchRange << 4);
  vo_clip_mv(&tlx, &tly, posX, posY, picW, picH, maxCuW, maxCuH);
  vo_clip_mv(&brx, &bry, posX, posY, picW, picH, maxCuW, maxCuH);
  *left   = vo_div_pow2(tlx, 4);
  *top    = vo_div_pow2(tly, 4);
  *right  = vo_div_pow2(brx, 4);
  *bottom = vo_div_pow2(bry, 4);
}

/* ------------------------------------------------------------------------------------------------
 * Integer full search
 * ---------------------------------------------------------------------------------------------- */

/* InterSearch::xPatternSearch — EncoderLib/InterSearch.cpp:3566-3608 */
void vo_pattern_search(const vo_job* j, int* mvx, int* mvy, uint64_t* sadOut)
{
  uint64_t  best     = UINT64_MAX;
  int       bx = 0, by = 0;
  const int subShift = vo_subshift(j->subShiftMode, j->w, j->h);
  for (int y = j->srTop; y <= j->srBottom; y++)
    for (int x = j->srLeft; x <= j->srRight; x++)
    {
      uint64_t sad = vo_sad(j->org, j->orgStride, j->refAtPU + y * j->refStride + x, j->refStride, j->w, j->h, subShift);
      sad += vo_mv_cost(j->lambdaMotion, vo_mv_bits(x, y, j->predQx, j->predQy, 2, j->imvShift));
      if (sad < best)
      {
        best = sad;
        bx   = x;
        by   = y;
      }
    }
  *mvx    = bx;
  *mvy    = by;
  *sadOut = best - vo_mv_cost(j->lambdaMotion, vo_mv_bits(bx, by, j->predQx, j->predQy, 2, j->imvShift));
}

/* ------------------------------------------------------------------------------------------------
 * Fractional refinement
 * ---------------------------------------------------------------------------------------------- */

static const int vo_refine_h[9][2] = { { 0, 0 },  { 0, -1 }, { 0, 1 },  { -1, 0 }, { 1, 0 },
                                       { -1, -1 }, { 1, -1 }, { -1, 1 }, { 1, 1 } }; /* InterSearch.cpp:59-70 */
static const int vo_refine_q[9][2] = { { 0, 0 },  { 0, -1 }, { 0, 1 },  { -1, -1 }, { 1, -1 },
                                       { -1, 0 }, { 1, 0 },  { -1, 1 }, { 1, 1 } }; /* :72-83 */

static uint64_t vo_dist(const vo_job* j, const vo_pel* cur, int curStride)
{
  return j->useHad ? vo_satd(j->org, j->orgStride, cur, curStride, j->w, j->h)
                   : vo_sad(j->org, j->orgStride, cur, curStride, j->w, j->h, 0);
}

#define VO_FB_STRIDE (VO_MAX_CU + 1)
#define VO_FB_SIZE ((VO_MAX_CU + 1 + 16) * (VO_MAX_CU + 8 + 8))
typedef struct
{
  vo_pel tmp[4][VO_FB_SIZE];    /* m_filteredBlockTmp[i][COMPONENT_Y] */
  vo_pel blk[4][4][VO_FB_SIZE]; /* m_filteredBlock[ver][hor][COMPONENT_Y] */
} vo_fb;

/* InterSearch::xExtDIFUpSamplingH — EncoderLib/InterSearch.cpp:5840-5882 */
static void vo_upsample_h(vo_fb* fb, const vo_pel* roi, int srcStride, int w, int h, int bd, int alt)
{
  const int     intStride = w + 1, dstStride = w + 1, half = 4;
  const vo_pel* srcPtr = roi - half * srcStride - 1;
  vo_filter_hor(0, srcPtr, srcStride, fb->tmp[0], intStride, w + 1, h + 8, 0, 0, bd, alt);
  vo_filter_hor(0, srcPtr, srcStride, fb->tmp[2], intStride, w + 1, h + 8, 8, 0, bd, alt);
  vo_filter_ver(0, fb->tmp[0] + half * intStride + 1, intStride, fb->blk[0][0], dstStride, w, h, 0, 0, 1, bd, alt);
  vo_filter_ver(0, fb->tmp[0] + (half - 1) * intStride + 1, intStride, fb->blk[2][0], dstStride, w, h + 1, 8, 0, 1, bd, alt);
  vo_filter_ver(0, fb->tmp[2] + half * intStride, intStride, fb->blk[0][2], dstStride, w + 1, h, 0, 0, 1, bd, alt);
  vo_filter_ver(0, fb->tmp[2] + (half - 1) * intStride, intStride, fb->blk[2][2], dstStride, w + 1, h + 1, 8, 0, 1, bd, alt);
}

/* InterSearch::xExtDIFUpSamplingQ — EncoderLib/InterSearch.cpp:5895-6050 */
static void vo_upsample_q(vo_fb* fb, const vo_pel* roi, int srcStride, int w, int h, int bd, int hx, int hy)
{
  const int     intStride = w + 1, dstStride = w + 1, half = 4;
  const int     extHeight = (hy == 0) ? h + 8 : h + 7;
  const vo_pel* srcPtr;
  vo_pel*       intPtr;

  srcPtr = roi - half * srcStride - 1; /* horizontal 1/4 */
  if (hy > 0) srcPtr += srcStride;
  if (hx >= 0) srcPtr += 1;
  vo_filter_hor(0, srcPtr, srcStride, fb->tmp[1], intStride, w, extHeight, 4, 0, bd, 0);

  srcPtr = roi - half * srcStride - 1; /* horizontal 3/4 */
  if (hy > 0) srcPtr += srcStride;
  if (hx > 0) srcPtr += 1;
  vo_filter_hor(0, srcPtr, srcStride, fb->tmp[3], intStride, w, extHeight, 12, 0, bd, 0);

  intPtr = fb->tmp[1] + (half - 1) * intStride; /* @1,1 */
  if (hy == 0) intPtr += intStride;
  vo_filter_ver(0, intPtr, intStride, fb->blk[1][1], dstStride, w, h, 4, 0, 1, bd, 0);

  intPtr = fb->tmp[1] + (half - 1) * intStride; /* @3,1 */
  vo_filter_ver(0, intPtr, intStride, fb->blk[3][1], dstStride, w, h, 12, 0, 1, bd, 0);

  if (hy != 0)
  {
    intPtr = fb->tmp[1] + (half - 1) * intStride; /* @2,1 */
    vo_filter_ver(0, intPtr, intStride, fb->blk[2][1], dstStride, w, h, 8, 0, 1, bd, 0);
    intPtr = fb->tmp[3] + (half - 1) * intStride; /* @2,3 */
    vo_filter_ver(0, intPtr, intStride, fb->blk[2][3], dstStride, w, h, 8, 0, 1, bd, 0);
  }
  else
  {
    intPtr = fb->tmp[1] + half * intStride; /* @0,1 */
    vo_filter_ver(0, intPtr, intStride, fb->blk[0][1], dstStride, w, h, 0, 0, 1, bd, 0);
    intPtr = fb->tmp[3] + half * intStride; /* @0,3 */
    vo_filter_ver(0, intPtr, intStride, fb->blk[0][3], dstStride, w, h, 0, 0, 1, bd, 0);
  }

  if (hx != 0)
  {
    intPtr = fb->tmp[2] + (half - 1) * intStride; /* @1,2 */
    if (hx > 0) intPtr += 1;
    if (hy >= 0) intPtr += intStride;
    vo_filter_ver(0, intPtr, intStride, fb->blk[1][2], dstStride, w, h, 4, 0, 1, bd, 0);
    intPtr = fb->tmp[2] + (half - 1) * intStride; /* @3,2 */
    if (hx > 0) intPtr += 1;
    if (hy > 0) intPtr += intStride;
    vo_filter_ver(0, intPtr, intStride, fb->blk[3][2], dstStride, w, h, 12, 0, 1, bd, 0);
  }
  else
  {
    intPtr = fb->tmp[0] + (half - 1) * intStride + 1; /* @1,0 */
    if (hy >= 0) intPtr += intStride;
    vo_filter_ver(0, intPtr, intStride, fb->blk[1][0], dstStride, w, h, 4, 0, 1, bd, 0);
    intPtr = fb->tmp[0] + (half - 1) * intStride + 1; /* @3,0 */
    if (hy > 0) intPtr += intStride;
    vo_filter_ver(0, intPtr, intStride, fb->blk[3][0], dstStride, w, h, 12, 0, 1, bd, 0);
  }

  intPtr = fb->tmp[3] + (half - 1) * intStride; /* @1,3 */
  if (hy == 0) intPtr += intStride;
  vo_filter_ver(0, intPtr, intStride, fb->blk[1][3], dstStride, w, h, 4, 0, 1, bd, 0);

  intPtr = fb->tmp[3] + (half - 1) * intStride; /* @3,3 */
  vo_filter_ver(0, intPtr, intStride, fb->blk[3][3], dstStride, w, h, 12, 0, 1, bd, 0);
}

/* InterSearch::xPatternRefinement — EncoderLib/InterSearch.cpp:707-761
 * (baseX,baseY) = baseRefMv, (*fx,*fy) in: MV for the rate term, out: best offset */
static uint64_t vo_refine_literal(const vo_job* j, vo_fb* fb, int baseX, int baseY, int iFrac, int* fx, int* fy)
{
  uint64_t best = UINT64_MAX;
  int      bestDir = 0;
  const int refStride = j->w + 1;
  const int (*tab)[2] = (iFrac == 2) ? vo_refine_h : vo_refine_q;
  const int costScale = (iFrac == 2) ? 1 : 0; /* set by the caller: InterSearch.cpp:4319, 4330 */
  for (int i = 0; i < 9; i++)
  {
    int horVal = (tab[i][0] + baseX) * iFrac;
    int verVal = (tab[i][1] + baseY) * iFrac;
    const vo_pel* p = fb->blk[verVal & 3][horVal & 3];
    if (horVal == 2 && (verVal & 1) == 0) p += 1;
    if ((horVal & 1) == 0 && verVal == 2) p += refStride;
    int      tx = tab[i][0] + *fx, ty = tab[i][1] + *fy;
    uint64_t d  = vo_dist(j, p, refStride);
    d += vo_mv_cost(j->lambdaMotion, vo_mv_bits(tx, ty, j->predQx, j->predQy, costScale, 0));
    if (d < best)
    {
      best    = d;
      bestDir = i;
    }
  }
  *fx = tab[bestDir][0];
  *fy = tab[bestDir][1];
  return best;
}

/* Body of InterSearch::xPatternSearchFracDIF — EncoderLib/InterSearch.cpp:4296-4338, literal buffers */
void vo_frac_literal(const vo_job* j, int mvx, int mvy, int* hx, int* hy, int* qx, int* qy, uint64_t* cost)
{
  const vo_pel* roi = j->refAtPU + mvx + mvy * j->refStride;
  *hx = *hy = *qx = *qy = 0;
  if (j->imvShift > 1) /* :4311-4317 */
  {
    *cost = vo_dist(j, roi, j->refStride) + vo_mv_cost(j->lambdaMotion, vo_mv_bits(mvx, mvy, j->predQx, j->predQy, 2, j->imvShift));
    return;
  }
  vo_fb* fb = (vo_fb*) malloc(sizeof(vo_fb));
  memset(fb, 0, sizeof(vo_fb));
  vo_upsample_h(fb, roi, j->refStride, j->w, j->h, j->bitDepth, j->useAltHpel);
  int fx = mvx << 1, fy = mvy << 1;
  *cost = vo_refine_literal(j, fb, 0, 0, 2, &fx, &fy);
  *hx   = fx;
  *hy   = fy;
  if (j->imvShift == 0)
  {
    vo_upsample_q(fb, roi, j->refStride, j->w, j->h, j->bitDepth, *hx, *hy);
    fx    = ((mvx << 1) + *hx) << 1;
    fy    = ((mvy << 1) + *hy) << 1;
    *cost = vo_refine_literal(j, fb, *hx << 1, *hy << 1, 1, &fx, &fy);
    *qx   = fx;
    *qy   = fy;
  }
  free(fb);
}

/* Direct form: the prediction block at quarter-pel offset (dqx,dqy) ∈ [-3,3]² from the integer MV is the
 * two-stage separable interpolation at integer base floor(d/4), phase d&3: horizontal pass
 * (isFirst,!isLast) to 14-bit intermediates over H+7 rows, vertical pass (!isFirst,isLast).  This is what
 * the m_filteredBlock planes of xExtDIFUpSamplingH/Q hold at the positions xPatternRefinement reads
 * (InterSearch.cpp:728-739); tests/test_oracle.py checks literal == direct. */
void vo_pred_qpel(const vo_job* j, int mvx, int mvy, int dqx, int dqy, int useAltHpel, vo_pel* dst, int dstStride)
{
  const int     ix = dqx >> 2, iy = dqy >> 2, px = dqx & 3, py = dqy & 3;
  const vo_pel* src = j->refAtPU + (mvx + ix) + (mvy + iy) * j->refStride;
  vo_pel        tmp[(VO_MAX_CU + 8) * VO_MAX_CU];
  const int     w = j->w, h = j->h;
  /* the 4x4 coefficient quirk (InterpolationFilter.cpp:786,869) keys on the (w,h) the reference passes:
   * never hit for inter CUs (no 4x4); the direct form uses the regular table. */
  const int16_t* ch = (px == 2 && useAltHpel) ? vo_luma_alt_hpel : vo_luma[px * 4];
  const int16_t* cv = (py == 2 && useAltHpel) ? vo_luma_alt_hpel : vo_luma[py * 4];
  if (px == 0)
    vo_filter_copy(1, 0, j->bitDepth, src - 3 * j->refStride, j->refStride, tmp, w, w, h + 7);
  else
    vo_filter(8, 0, 1, 0, j->bitDepth, src - 3 * j->refStride, j->refStride, tmp, w, w, h + 7, ch);
  if (py == 0)
    vo_filter_copy(0, 1, j->bitDepth, tmp + 3 * w, w, dst, dstStride, w, h);
  else
    vo_filter(8, 1, 0, 1, j->bitDepth, tmp + 3 * w, w, dst, dstStride, w, h, cv);
}

void vo_frac_direct(const vo_job* j, int mvx, int mvy, int* hx, int* hy, int* qx, int* qy, uint64_t* cost)
{
  const vo_pel* roi = j->refAtPU + mvx + mvy * j->refStride;
  vo_pel        pred[VO_MAX_CU * VO_MAX_CU];
  *hx = *hy = *qx = *qy = 0;
  if (j->imvShift > 1)
  {
    *cost = vo_dist(j, roi, j->refStride) + vo_mv_cost(j->lambdaMotion, vo_mv_bits(mvx, mvy, j->predQx, j->predQy, 2, j->imvShift));
    return;
  }
  uint64_t best = UINT64_MAX;
  int      dir  = 0;
  for (int i = 0; i < 9; i++) /* half-pel, cost scale 1 */
  {
    vo_pred_qpel(j, mvx, mvy, vo_refine_h[i][0] * 2, vo_refine_h[i][1] * 2, j->useAltHpel, pred, j->w);
    uint64_t d = vo_dist(j, pred, j->w);
    d += vo_mv_cost(j->lambdaMotion,
                    vo_mv_bits((mvx << 1) + vo_refine_h[i][0], (mvy << 1) + vo_refine_h[i][1], j->predQx, j->predQy, 1, 0));
    if (d < best) best = d, dir = i;
  }
  *hx   = vo_refine_h[dir][0];
  *hy   = vo_refine_h[dir][1];
  *cost = best;
  if (j->imvShift == 0)
  {
    best = UINT64_MAX;
    dir  = 0;
    for (int i = 0; i < 9; i++) /* quarter-pel, cost scale 0 */
    {
      int dqx = *hx * 2 + vo_refine_q[i][0], dqy = *hy * 2 + vo_refine_q[i][1];
      vo_pred_qpel(j, mvx, mvy, dqx, dqy, 0, pred, j->w);
      uint64_t d = vo_dist(j, pred, j->w);
      d += vo_mv_cost(j->lambdaMotion, vo_mv_bits((mvx << 2) + dqx, (mvy << 2) + dqy, j->predQx, j->predQy, 0, 0));
      if (d < best) best = d, dir = i;
    }
    *qx   = vo_refine_q[dir][0];
    *qy   = vo_refine_q[dir][1];
    *cost = best;
  }
}

/* xPatternSearch followed by the xPatternSearchFracDIF body, as xMotionEstimation chains them
 * (EncoderLib/InterSearch.cpp:3432, 3476) */
void vo_search(const vo_job* j, vo_result* r, int literal)
{
  vo_pattern_search(j, &r->mvX, &r->mvY, &r->intSad);
  r->halfX = r->halfY = r->qterX = r->qterY = 0;
  r->fracCost = r->intSad;
  if (!j->doFrac) return;
  if (literal)
    vo_frac_literal(j, r->mvX, r->mvY, &r->halfX, &r->halfY, &r->qterX, &r->qterY, &r->fracCost);
  else
    vo_frac_direct(j, r->mvX, r->mvY, &r->halfX, &r->halfY, &r->qterX, &r->qterY, &r->fracCost);
}

double vo_search_batch(const vo_job* jobs, vo_result* res, int n, int literal)
{
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < n; i++) vo_search(&jobs[i], &res[i], literal);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

/* ---- motion compensation (SURVEY.md §8f "next": the caller-side consumer of the found MVs) ------------------------- */

/* InterPrediction::xPredInterBlk — CommonLib/InterPrediction.cpp:660-830, the plain path (no RPR, wrap-around, BDOF
 * padding, DMVR or bilinear).  refAtBlk: component plane at the block position; mv in 1/16 luma sample, so the
 * fractional part has 4 bits for luma and 5 for 4:2:0 chroma (:675-676, :697-698).  bi != 0 leaves the 14-bit
 * intermediates for the weighted average (rndRes = !bi, :673). */
void vo_mc_block(int comp, const vo_pel* refAtBlk, int refStride, int w, int h, int mvX, int mvY, int bi, int bd,
                 int useAltHpel, vo_pel* dst, int dstStride)
{
  const int     shift = 4 + (comp ? 1 : 0);
  const int     xFrac = mvX & ((1 << shift) - 1), yFrac = mvY & ((1 << shift) - 1);
  const vo_pel* src   = refAtBlk + (ptrdiff_t) (mvY >> shift) * refStride + (mvX >> shift);
  const int     rnd   = !bi;
  if (yFrac == 0)
    vo_filter_hor(comp, src, refStride, dst, dstStride, w, h, xFrac, rnd, bd, useAltHpel);
  else if (xFrac == 0)
    vo_filter_ver(comp, src, refStride, dst, dstStride, w, h, yFrac, 1, rnd, bd, useAltHpel);
  else
  {
    const int taps = comp ? 4 : VO_NTAPS_LUMA, above = (taps >> 1) - 1;
    vo_pel    tmp[(VO_MAX_CU + 8) * VO_MAX_CU];
    vo_filter_hor(comp, src - above * refStride, refStride, tmp, w, w, h + taps - 1, xFrac, 0, bd, useAltHpel);
    vo_filter_ver(comp, tmp + above * w, w, dst, dstStride, w, h, yFrac, 0, rnd, bd, useAltHpel);
  }
}

/* AreaBuf<Pel>::addAvg — CommonLib/Buffer.cpp:467-507 (the default bi-prediction average of two bi=1 predictions) */
void vo_add_avg(const vo_pel* s0, const vo_pel* s1, vo_pel* dst, int n, int bd)
{
  const int shift  = ((VO_IF_INTERNAL_PREC - bd) > 2 ? (VO_IF_INTERNAL_PREC - bd) : 2) + 1;
  const int offset = (1 << (shift - 1)) + 2 * VO_IF_INTERNAL_OFFS;
  const int maxv   = (1 << bd) - 1;
  for (int i = 0; i < n; i++)
  {
    int v  = (s0[i] + s1[i] + offset) >> shift;
    dst[i] = (vo_pel) (v < 0 ? 0 : (v > maxv ? maxv : v));
  }
}

/* AreaBuf<Pel>::addWeightedAvg — CommonLib/Buffer.cpp:365-396: the bi-prediction average under BCW weights
 * (w1 = g_BcwWeights[bcwIdx], w0 = 8 - w1; Rom.cpp:188-200) of two bi=1 predictions */
void vo_add_weighted_avg(const vo_pel* s0, const vo_pel* s1, vo_pel* dst, int n, int bd, int bcwIdx)
{
  static const int bcwWeights[5] = { -2, 3, 4, 5, 10 };
  const int w1 = bcwWeights[bcwIdx], w0 = 8 - w1;
  const int shift  = ((VO_IF_INTERNAL_PREC - bd) > 2 ? (VO_IF_INTERNAL_PREC - bd) : 2) + 3;
  const int offset = (1 << (shift - 1)) + (VO_IF_INTERNAL_OFFS << 3);
  const int maxv   = (1 << bd) - 1;
  for (int i = 0; i < n; i++)
  {
    int v  = (s0[i] * w0 + s1[i] * w1 + offset) >> shift;
    dst[i] = (vo_pel) (v < 0 ? 0 : (v > maxv ? maxv : v));
  }
}

/* AreaBuf<T>::removeHighFreq — CommonLib/Buffer.h:474-517: the bi-prediction search target 2*org - otherPred */
void vo_remove_high_freq(vo_pel* dst, const vo_pel* src, int n, int clip, int bd)
{
  const int maxv = (1 << bd) - 1;
  for (int i = 0; i < n; i++)
  {
    int v = 2 * dst[i] - src[i];
    if (clip) v = v < 0 ? 0 : (v > maxv ? maxv : v);
    dst[i] = (vo_pel) v;
  }
}

/* Tail of xMotionEstimation — EncoderLib/InterSearch.cpp:3477-3484 */
void vo_me_finish(const vo_job* j, const vo_result* r, double fWeight, uint32_t bitsIn, int* mvQx, int* mvQy,
                  uint32_t* bitsOut, uint64_t* costOut)
{
  int      mx = (r->mvX << 2) + (r->halfX << 1) + r->qterX;
  int      my = (r->mvY << 2) + (r->halfY << 1) + r->qterY;
  uint32_t mvBits = vo_mv_bits(mx, my, j->predQx, j->predQy, 0, j->imvShift);
  uint32_t bits   = bitsIn + mvBits;
  *mvQx    = mx;
  *mvQy    = my;
  *bitsOut = bits;
  *costOut = (uint64_t) (floor(fWeight * ((double) r->fracCost - (double) vo_mv_cost(j->lambdaMotion, mvBits))) +
                         (double) vo_mv_cost(j->lambdaMotion, bits));
}

/* ------------------------------------------------------------------------------------------------
 * Integer / 4-pel AMVR refinement
 * ---------------------------------------------------------------------------------------------- */

/* Mv::changePrecision — CommonLib/Mv.h:183-197: left shift, or right shift rounding to nearest, ties toward zero */
static int vo_change_prec(int v, int shift)
{
  if (shift >= 0) return v * (1 << shift);
  {
    const int rs = -shift, off = 1 << (rs - 1);
    return v >= 0 ? (v + off - 1) >> rs : (v + off) >> rs;
  }
}

/* InterSearch::xPatternSearchIntRefine — EncoderLib/InterSearch.cpp:4172-4282 (no MCTS constraint).
 * j supplies the pattern, the reference plane at the PU position, bit depth, useHad and lambda; subShift is 0
 * (setDistParam(..., 0, 1, useHad), :4179). */
void vo_int_refine(const vo_job* j, vo_int_refine_io* io)
{
  static const int testPos[9][2] = { { 0, 0 }, { -1, -1 }, { -1, 0 }, { -1, 1 }, { 0, -1 }, { 0, 1 }, { 1, -1 }, { 1, 0 }, { 1, 1 } };
  /* Mv::m_amvrPrecision (Mv.cpp:41): imv 1 -> MV_PRECISION_INT (2), imv 2 -> MV_PRECISION_4PEL (0); internal = 6 */
  const int prec  = io->imv == 1 ? 2 : 0;
  const int down  = prec - 6, up = 6 - prec;
  uint64_t  dist, satd = 0, bestDist = UINT64_MAX;
  uint32_t  bits = io->bits - io->mvpIdxBits[io->mvpIdx]; /* :4188 */
  int       bestX = io->mvX, bestY = io->mvY, bestBits = 0, bestIdx = io->mvpIdx;
  int       baseX[2], baseY[2], testX[2] = { 0, 0 }, testY[2] = { 0, 0 };
  int       pos, i;
  for (i = 0; i < 2; i++) /* cBaseMvd[i] = roundTransPrecInternal2Amvr(rcMv - mvCand[i]), :4198-4204 */
  {
    baseX[i] = vo_change_prec(vo_change_prec(io->mvX - io->candX[i], down), up);
    baseY[i] = vo_change_prec(vo_change_prec(io->mvY - io->candY[i], down), up);
  }
  for (pos = 0; pos < 9; pos++)
    for (i = 0; i < io->numCand; i++)
    {
      int      mvBits, px, py, mx, my;
      testX[i] = vo_change_prec(testPos[pos][0], up) + baseX[i] + io->candX[i]; /* :4214-4217 */
      testY[i] = vo_change_prec(testPos[pos][1], up) + baseY[i] + io->candY[i];
      if (i == 0 || testX[0] != testX[1] || testY[0] != testY[1])
      {
        int cx = testX[i], cy = testY[i];
        vo_clip_mv(&cx, &cy, io->posX, io->posY, io->picW, io->picH, io->maxCuW, io->maxCuH); /* :4237 */
        {
          const vo_pel* cur = j->refAtPU + (ptrdiff_t) j->refStride * (cy >> 4) + (cx >> 4);
          dist = satd = (uint64_t) ((double) vo_dist(j, cur, j->refStride) * io->fWeight); /* :4240 */
        }
      }
      else
        dist = satd;
      mvBits = (int) io->mvpIdxBits[i];
      px     = vo_change_prec(io->candX[i], down);
      py     = vo_change_prec(io->candY[i], down);
      mx     = vo_change_prec(testX[i], down);
      my     = vo_change_prec(testY[i], down);
      mvBits += (int) vo_mv_bits(mx, my, px, py, 0, 0);
      dist += vo_mv_cost(j->lambdaMotion, (uint32_t) mvBits);
      if (dist < bestDist)
      {
        bestDist = dist;
        bestX    = testX[i];
        bestY    = testY[i];
        bestIdx  = i;
        bestBits = mvBits;
      }
    }
  if (bestDist == UINT64_MAX)
  {
    io->bits = bits;
    io->cost = UINT64_MAX;
    return;
  }
  io->mvX    = bestX;
  io->mvY    = bestY;
  io->mvpIdx = bestIdx;
  bits += (uint32_t) bestBits;
  io->bits = bits;
  io->cost = bestDist - vo_mv_cost(j->lambdaMotion, (uint32_t) bestBits) + vo_mv_cost(j->lambdaMotion, bits); /* :4276 */
}

/* ------------------------------------------------------------------------------------------------
 * TZ search
 * ---------------------------------------------------------------------------------------------- */

typedef struct
{
  const vo_job* j;
  int           subShift;
  int           l, r, t, b; /* cStruct.searchRange */
  uint64_t      bestSad;
  int           bestX, bestY;
  unsigned      bestDistance, bestRound;
  int           pointNr;
  int           probes;
} vo_tz;

/* xTZSearchHelp, subShiftMode == 1 branch — InterSearch.cpp:340-391 (the selective search's staged SAD): the rows
 * 0, 2^S, 2*2^S ... first, scaled up by 2^S as an estimate of the whole SAD; then the rows half-way between those
 * already summed, stage by stage, each time testing the scaled partial sum against the best cost.  Only a probe that
 * survives every stage (and so has its exact SAD) can become the best.  The tests are estimates: a probe whose exact
 * cost is below the best can be rejected, so the outcome depends on the best cost at the time of the probe. */
static void vo_tz_probe_staged(vo_tz* s, int x, int y, int pointNr, unsigned dist)
{
  const vo_job*  j       = s->j;
  const vo_pel*  cur     = j->refAtPU + (ptrdiff_t) y * j->refStride + x;
  const uint64_t bitCost = vo_mv_cost(j->lambdaMotion, vo_mv_bits(x, y, j->predQx, j->predQy, 2, j->imvShift));
  int            sub     = s->subShift;
  uint64_t       tmp, sad = 0;
  if (!(bitCost < s->bestSad)) return; /* :346 */
  s->probes++;
  tmp = vo_sad(j->org, j->orgStride, cur, j->refStride, j->w, j->h, sub); /* = (sum over rows 0 mod 2^sub) << sub */
  if (!(tmp + bitCost < s->bestSad)) return; /* :350 */
  sad += tmp >> sub;
  while (sub > 0) /* :357-371 */
  {
    const int isub = sub - 1;
    tmp = vo_sad(j->org + ((ptrdiff_t) j->orgStride << isub), j->orgStride, cur + ((ptrdiff_t) j->refStride << isub), j->refStride,
                 j->w, j->h, sub); /* distFunc walks h >> sub rows from the offset row: rows 2^isub mod 2^sub */
    sad += tmp >> sub;
    if ((sad << isub) + bitCost > s->bestSad) break;
    sub--;
  }
  if (sub == 0)
  {
    sad += bitCost;
    if (sad < s->bestSad)
    {
      s->bestSad      = sad;
      s->bestX        = x;
      s->bestY        = y;
      s->bestDistance = dist;
      s->bestRound    = 0;
      s->pointNr      = pointNr;
    }
  }
}

/* xTZSearchHelp — InterSearch.cpp:330-417 */
static void vo_tz_probe(vo_tz* s, int x, int y, int pointNr, unsigned dist)
{
  const vo_job* j = s->j;
  uint64_t      sad;
  if (j->subShiftMode == 1)
  {
    vo_tz_probe_staged(s, x, y, pointNr, dist);
    return;
  }
  sad = vo_sad(j->org, j->orgStride, j->refAtPU + (ptrdiff_t) y * j->refStride + x, j->refStride, j->w, j->h, s->subShift);
  s->probes++;
  if (sad < s->bestSad)
  {
    sad += vo_mv_cost(j->lambdaMotion, vo_mv_bits(x, y, j->predQx, j->predQy, 2, j->imvShift));
    if (sad < s->bestSad)
    {
      s->bestSad      = sad;
      s->bestX        = x;
      s->bestY        = y;
      s->bestDistance = dist;
      s->bestRound    = 0;
      s->pointNr      = pointNr;
    }
  }
}

/* One point of a diamond: a coordinate that moved away from the start is checked against the window on its side,
 * a coordinate equal to the start's is not checked — the rule every branch of xTZ8PointDiamondSearch follows. */
static void vo_tz_point(vo_tz* s, int sx, int sy, int dx, int dy, int pointNr, unsigned dist)
{
  const int x = sx + dx, y = sy + dy;
  if ((dx < 0 && x < s->l) || (dx > 0 && x > s->r) || (dy < 0 && y < s->t) || (dy > 0 && y > s->b)) return;
  vo_tz_probe(s, x, y, pointNr, dist);
}

/* xTZ8PointDiamondSearch — InterSearch.cpp:503-705 (probe order and point numbers as there) */
static void vo_tz_diamond(vo_tz* s, int sx, int sy, int d, int cornersAtDist1)
{
  s->bestRound += 1;
  if (d == 1)
  {
    if (cornersAtDist1)
    {
      if (sy - 1 >= s->t)
      {
        vo_tz_point(s, sx, sy, -1, -1, 1, 1);
        vo_tz_point(s, sx, sy, 0, -1, 2, 1);
        vo_tz_point(s, sx, sy, 1, -1, 3, 1);
      }
    }
    else
      vo_tz_point(s, sx, sy, 0, -1, 2, 1);
    vo_tz_point(s, sx, sy, -1, 0, 4, 1);
    vo_tz_point(s, sx, sy, 1, 0, 5, 1);
    if (cornersAtDist1)
    {
      if (sy + 1 <= s->b)
      {
        vo_tz_point(s, sx, sy, -1, 1, 6, 1);
        vo_tz_point(s, sx, sy, 0, 1, 7, 1);
        vo_tz_point(s, sx, sy, 1, 1, 8, 1);
      }
    }
    else
      vo_tz_point(s, sx, sy, 0, 1, 7, 1);
  }
  else if (d <= 8)
  {
    const int h = d >> 1;
    vo_tz_point(s, sx, sy, 0, -d, 2, d);
    vo_tz_point(s, sx, sy, -h, -h, 1, h);
    vo_tz_point(s, sx, sy, h, -h, 3, h);
    vo_tz_point(s, sx, sy, -d, 0, 4, d);
    vo_tz_point(s, sx, sy, d, 0, 5, d);
    vo_tz_point(s, sx, sy, -h, h, 6, h);
    vo_tz_point(s, sx, sy, h, h, 8, h);
    vo_tz_point(s, sx, sy, 0, d, 7, d);
  }
  else
  {
    int i;
    vo_tz_point(s, sx, sy, 0, -d, 0, d);
    vo_tz_point(s, sx, sy, -d, 0, 0, d);
    vo_tz_point(s, sx, sy, d, 0, 0, d);
    vo_tz_point(s, sx, sy, 0, d, 0, d);
    for (i = 1; i < 4; i++)
    {
      const int q = (d >> 2) * i;
      vo_tz_point(s, sx, sy, -q, -d + q, 0, d);
      vo_tz_point(s, sx, sy, q, -d + q, 0, d);
      vo_tz_point(s, sx, sy, -q, d - q, 0, d);
      vo_tz_point(s, sx, sy, q, d - q, 0, d);
    }
  }
}

/* xTZ2PointSearch — InterSearch.cpp:420-446 */
static void vo_tz_two_points(vo_tz* s)
{
  static const int xo[2][9] = { { 0, -1, -1, 0, -1, +1, -1, -1, +1 }, { 0, 0, +1, +1, -1, +1, 0, +1, 0 } };
  static const int yo[2][9] = { { 0, 0, -1, -1, +1, -1, 0, +1, 0 }, { 0, -1, -1, 0, -1, +1, +1, +1, +1 } };
  int              k;
  const int        bx = s->bestX, by = s->bestY, pn = s->pointNr; /* both points are derived before the first probe */
  for (k = 0; k < 2; k++)
  {
    const int x = bx + xo[k][pn], y = by + yo[k][pn];
    if (x >= s->l && x <= s->r && y >= s->t && y <= s->b) vo_tz_probe(s, x, y, 0, 2);
  }
}

/* clipMv + changePrecision(INTERNAL -> QUARTER) + divideByPowerOf2(2) — InterSearch.cpp:3682-3683 */
static void vo_tz_to_int(const vo_tz_params* p, int* x, int* y)
{
  vo_clip_mv(x, y, p->posX, p->posY, p->picW, p->picH, p->maxCuW, p->maxCuH);
  *x = vo_div_pow2(vo_change_prec(*x, -2), 2);
  *y = vo_div_pow2(vo_change_prec(*y, -2), 2);
}

/* History MVs of xTZSearch (:3734-3765) and xTZSearchSelective (:4042-4074): duplicates of a newer entry are skipped;
 * the distortion is distFunc's own (sub-sampled rows scaled up, also for subShiftMode 1); only position and cost change */
static void vo_tz_seeds(vo_tz* s, const vo_tz_params* p)
{
  const vo_job* j = s->j;
  int           i, k;
  for (i = 0; i < p->nSeeds; i++)
  {
    int      x = p->seedX[i], y = p->seedY[i];
    uint64_t sad;
    for (k = 0; k < i; k++)
      if (p->seedX[k] == x && p->seedY[k] == y) break;
    if (k < i) continue;
    vo_clip_mv(&x, &y, p->posX, p->posY, p->picW, p->picH, p->maxCuW, p->maxCuH);
    x   = vo_change_prec(x, -4);
    y   = vo_change_prec(y, -4);
    sad = vo_sad(j->org, j->orgStride, j->refAtPU + (ptrdiff_t) y * j->refStride + x, j->refStride, j->w, j->h, s->subShift);
    s->probes++;
    sad += vo_mv_cost(j->lambdaMotion, vo_mv_bits(x, y, j->predQx, j->predQy, 2, j->imvShift));
    if (sad < s->bestSad)
    {
      s->bestSad = sad;
      s->bestX   = x;
      s->bestY   = y;
    }
  }
}

/* InterSearch::xTZSearchSelective — EncoderLib/InterSearch.cpp:3979-4170 (FastSearch=2, MESEARCH_SELECTIVE; no hash ME) */
static void vo_tz_search_selective(const vo_job* j, const vo_tz_params* p, int* mvx, int* mvy, uint64_t* sadOut, int* nProbes)
{
  const int range = p->searchRange, rangeInitial = p->searchRange >> 2, step = 4, distThresh = 8;
  vo_tz     s;
  int       sx, sy, d, x, y, bx, by, l, r, t, b;
  memset(&s, 0, sizeof(s));
  s.j        = j;
  s.subShift = vo_subshift(j->subShiftMode, j->w, j->h);
  s.bestSad  = UINT64_MAX;

  sx = p->startX;
  sy = p->startY;
  vo_tz_to_int(p, &sx, &sy);       /* :4003-4005 */
  vo_tz_probe(&s, sx, sy, 0, 0);   /* :4017 */
  vo_tz_probe(&s, 0, 0, 0, 0);     /* :4020-4023, bTestZeroVector: unconditional here */
  if (p->hasInt2Nx2N)              /* :4027-4038: no duplicate test either */
  {
    int ix = p->int2Nx2NX * 16, iy = p->int2Nx2NY * 16;
    vo_tz_to_int(p, &ix, &iy);
    vo_tz_probe(&s, ix, iy, 0, 0);
  }
  vo_tz_seeds(&s, p);              /* :4040-4074 */
  /* :4076-4081 — the reference shifts the integer position by 2, not by MV_FRACTIONAL_BITS_INTERNAL, before handing it to
   * xSetSearchRange as a 1/16-sample vector: the window is centred on a quarter of the best start point.  Restated as is. */
  vo_set_search_range(s.bestX * 4, s.bestY * 4, p->posX, p->posY, p->picW, p->picH, p->maxCuW, p->maxCuH, range, &s.l, &s.r,
                      &s.t, &s.b);

  bx = s.bestX; /* initial search: a grid of step 4 around the best start point, each point with its two smallest diamonds, :4104-4120 */
  by = s.bestY;
  l  = bx - rangeInitial > s.l ? bx - rangeInitial : s.l;
  t  = by - rangeInitial > s.t ? by - rangeInitial : s.t;
  r  = bx + rangeInitial < s.r ? bx + rangeInitial : s.r;
  b  = by + rangeInitial < s.b ? by + rangeInitial : s.b;
  for (y = t; y <= b; y += step)
    for (x = l; x <= r; x += step)
    {
      vo_tz_probe(&s, x, y, 0, 0);
      vo_tz_diamond(&s, x, y, 1, 0);
      vo_tz_diamond(&s, x, y, 2, 0);
    }

  if (abs(s.bestX - bx) > distThresh || abs(s.bestY - by) > distThresh) /* far from the predictors: full search, :4122-4134 */
  {
    for (y = s.t; y <= s.b; y++)
      for (x = s.l; x <= s.r; x++) vo_tz_probe(&s, x, y, 0, 1);
  }
  else /* star refinement without a stop criterion, :4136-4165 */
    while (s.bestDistance > 0)
    {
      sx             = s.bestX;
      sy             = s.bestY;
      s.bestDistance = 0;
      s.pointNr      = 0;
      for (d = 1; d < range + 1; d *= 2) vo_tz_diamond(&s, sx, sy, d, 0);
      if (s.bestDistance == 1)
      {
        s.bestDistance = 0;
        if (s.pointNr != 0) vo_tz_two_points(&s);
      }
    }
  *mvx    = s.bestX;
  *mvy    = s.bestY;
  *sadOut = s.bestSad - vo_mv_cost(j->lambdaMotion, vo_mv_bits(s.bestX, s.bestY, j->predQx, j->predQy, 2, j->imvShift));
  if (nProbes) *nProbes = s.probes;
}

/* InterSearch::xTZSearch — EncoderLib/InterSearch.cpp:3640-3974 */
void vo_tz_search(const vo_job* j, const vo_tz_params* p, int* mvx, int* mvy, uint64_t* sadOut, int* nProbes)
{
  const int raster      = p->fast ? 8 : 5;               /* iRaster */
  const int firstRounds = 3;                             /* uiFirstSearchRounds (both settings) */
  const int range       = p->searchRange;
  vo_tz     s;
  int       sx, sy, d, bestIsZero;
  if (p->selective)
  {
    vo_tz_search_selective(j, p, mvx, mvy, sadOut, nProbes);
    return;
  }
  memset(&s, 0, sizeof(s));
  s.j        = j;
  s.subShift = vo_subshift(j->subShiftMode, j->w, j->h);
  s.bestSad  = UINT64_MAX;

  sx = p->startX;
  sy = p->startY;
  vo_tz_to_int(p, &sx, &sy);
  vo_tz_probe(&s, sx, sy, 0, 0); /* :3695 */
  if (!p->fast && (sx != 0 || sy != 0) && (s.bestX != 0 || s.bestY != 0)) vo_tz_probe(&s, 0, 0, 0, 0); /* :3698-3707 */

  if (p->hasInt2Nx2N) /* :3711-3732 */
  {
    int ix = p->int2Nx2NX * 16, iy = p->int2Nx2NY * 16;
    vo_tz_to_int(p, &ix, &iy);
    if ((sx != ix || sy != iy) && (ix != s.bestX || iy != s.bestY)) vo_tz_probe(&s, ix, iy, 0, 0);
  }

  vo_tz_seeds(&s, p); /* :3734-3765 */

  vo_set_search_range(s.bestX * 16, s.bestY * 16, p->posX, p->posY, p->picW, p->picH, p->maxCuW, p->maxCuH,
                      range >> (p->fast ? 1 : 0), &s.l, &s.r, &s.t, &s.b); /* :3767-3772 */

  sx         = s.bestX;
  sy         = s.bestY;
  bestIsZero = s.bestX == 0 && s.bestY == 0;
  for (d = 1; d <= range; d *= 2) /* first search, :3803-3818 */
  {
    vo_tz_diamond(&s, sx, sy, d, p->extended);
    if (p->firstSearchStop && s.bestRound >= (unsigned) firstRounds) break;
  }
  if (p->extended && !bestIsZero) /* zero neighbourhood with half the range, :3841-3855 (bNewZeroNeighbourhoodTest) */
    for (d = 1; d <= (range >> 1); d *= 2) vo_tz_diamond(&s, 0, 0, d, 0);

  if (s.bestDistance == 1) /* :3858-3863 */
  {
    s.bestDistance = 0;
    vo_tz_two_points(&s);
  }

  if (p->extended) /* adaptive raster, :3865-3885 */
  {
    int win = raster, l = s.l, r = s.r, t = s.t, b = s.b, x, y;
    if (!((int) s.bestDistance >= raster))
    {
      win++;
      l /= 2;
      r /= 2;
      t /= 2;
      b /= 2;
    }
    s.bestDistance = (unsigned) win;
    for (y = t; y <= b; y += win)
      for (x = l; x <= r; x += win) vo_tz_probe(&s, x, y, 0, (unsigned) win);
  }
  else if ((int) s.bestDistance >= raster) /* :3886-3899 */
  {
    int x, y;
    s.bestDistance = (unsigned) raster;
    for (y = s.t; y <= s.b; y += raster)
      for (x = s.l; x <= s.r; x += raster) vo_tz_probe(&s, x, y, 0, (unsigned) raster);
  }

  while (s.bestDistance > 0) /* star refinement, :3932-3967 */
  {
    sx             = s.bestX;
    sy             = s.bestY;
    s.bestDistance = 0;
    s.pointNr      = 0;
    for (d = 1; d < range + 1; d *= 2)
    {
      vo_tz_diamond(&s, sx, sy, d, p->extended);
      if (p->fast && s.bestRound >= 2) break;
    }
    if (s.bestDistance == 1)
    {
      s.bestDistance = 0;
      if (s.pointNr != 0) vo_tz_two_points(&s);
    }
  }
  *mvx    = s.bestX;
  *mvy    = s.bestY;
  *sadOut = s.bestSad - vo_mv_cost(j->lambdaMotion, vo_mv_bits(s.bestX, s.bestY, j->predQx, j->predQy, 2, j->imvShift));
  if (nProbes) *nProbes = s.probes;
}

/* EncTemporalFilter::bilateralFilter, the weighting of one component — EncoderLib/EncTemporalFilter.cpp:568-621.
 * corrected[i]: the i-th neighbouring picture after applyMotion (vo_mctf_apply_motion), same geometry as org.
 * The weight of a reference sample depends only on the integer difference refVal - orgVal, the component class and the
 * POC distance class: w = (weightScaling * refStrength) * exp(-(diff * 1024 / 2^bd)^2 / (2 sigma^2)); the sums run in
 * reference order in double precision, without contraction (the reference is built without FMA), one division, round(). */
void vo_mctf_bilateral(const vo_pel* org, int orgStride, const vo_pel* const* corrected, int corrStride, const int* origOffset,
                       int numRefs, int w, int h, int isChroma, int qp, double overallStrength, int bitDepth, vo_pel* dst, int dstStride)
{
  static const double refStrengths[3][2] = { { 0.85, 0.60 }, { 1.20, 1.00 }, { 0.30, 0.30 } }; /* m_refStrengths, :72-78 */
  const int           range = 2;                                                             /* m_range */
  const int           row   = numRefs == range * 2 ? 0 : (numRefs == range ? 1 : 2);
  const double lumaSigmaSq = (qp - 10.0) * (qp - 10.0) * 9.0, chromaSigmaSq = 30 * 30;      /* m_sigmaZeroPoint, m_sigmaMultiplier */
  const double sigmaSq = isChroma ? chromaSigmaSq : lumaSigmaSq;
  const double weightScaling = overallStrength * (isChroma ? 0.55 : 0.4);                    /* m_chromaFactor */
  const int    maxv = (1 << bitDepth) - 1;
  const double bitDepthDiffWeighting = 1024.0 / (maxv + 1);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
    {
      const int       orgVal = org[(ptrdiff_t) y * orgStride + x];
      volatile double sum = 1.0, newVal = (double) orgVal; /* volatile: every product and sum rounded to double on its own */
      for (int i = 0; i < numRefs; i++)
      {
        const int    refVal = corrected[i][(ptrdiff_t) y * corrStride + x];
        double       diff   = (double) (refVal - orgVal);
        int          a      = origOffset[i] < 0 ? -origOffset[i] : origOffset[i];
        const int    index  = a - 1 < 1 ? a - 1 : 1;
        double       diffSq, weight;
        volatile double prod;
        diff *= bitDepthDiffWeighting;
        diffSq = diff * diff;
        weight = weightScaling * refStrengths[row][index] * exp(-diffSq / (2 * sigmaSq));
        prod   = weight * refVal;
        newVal = newVal + prod;
        sum    = sum + weight;
      }
      {
        const double q = newVal / sum;
        int          v = (int) (vo_pel) round(q);
        dst[(ptrdiff_t) y * dstStride + x] = (vo_pel) (v < 0 ? 0 : (v > maxv ? maxv : v));
      }
    }
}

/* The weights of vo_mctf_bilateral as a table over |refVal - orgVal| (they depend on nothing else within a component and a
 * POC-distance class): what a device implementation receives from the host instead of evaluating exp() itself, so that
 * its double-precision sums are those of the reference bit for bit.  table: 1 << bitDepth entries. */
void vo_mctf_bilateral_weights(int isChroma, int qp, double overallStrength, int bitDepth, int numRefs, int index, double* table)
{
  static const double refStrengths[3][2] = { { 0.85, 0.60 }, { 1.20, 1.00 }, { 0.30, 0.30 } };
  const int    row = numRefs == 4 ? 0 : (numRefs == 2 ? 1 : 2);
  const double sigmaSq = isChroma ? 30 * 30 : (qp - 10.0) * (qp - 10.0) * 9.0;
  const double weightScaling = overallStrength * (isChroma ? 0.55 : 0.4);
  const double bitDepthDiffWeighting = 1024.0 / (1 << bitDepth);
  for (int d = 0; d < (1 << bitDepth); d++)
  {
    double diff = (double) d;
    double diffSq;
    diff *= bitDepthDiffWeighting;
    diffSq   = diff * diff; /* (-d)^2 == d^2 exactly */
    table[d] = weightScaling * refStrengths[row][index] * exp(-diffSq / (2 * sigmaSq));
  }
}

/* ------------------------------------------------------------------------------------------------
 * Symmetric MVD search (SMVD)
 * ---------------------------------------------------------------------------------------------- */

/* AreaBuf<T>::removeWeightHighFreq — CommonLib/Buffer.h:418-472: the bi-prediction search target under a BCW weight,
 * (org * 8 - other * (8 - w)) / w in 16-bit fixed point (g_BcwWeightBase = 8, Rom.cpp:188-190) */
void vo_remove_weight_high_freq(vo_pel* dst, const vo_pel* src, int n, int clip, int bd, int bcwWeight)
{
  const int normalizer = ((1 << 16) + (bcwWeight > 0 ? (bcwWeight >> 1) : -(bcwWeight >> 1))) / bcwWeight;
  const int weight0 = normalizer * 8, weight1 = (8 - bcwWeight) * normalizer, maxv = (1 << bd) - 1;
  for (int i = 0; i < n; i++)
  {
    int v = (dst[i] * weight0 - src[i] * weight1 + (1 << 15)) >> 16;
    if (clip) v = v < 0 ? 0 : (v > maxv ? maxv : v);
    dst[i] = (vo_pel) v;
  }
}

/* InterSearch::xGetSymmetricCost — EncoderLib/InterSearch.cpp:4341-4391: both predictions with the
 * 8-tap filter at the clipped MVs (an integer MV reads the picture directly, which is what the filter's copy gives),
 * 2*org - predA (removeHighFreq; removeWeightHighFreq under a BCW weight), SATD or SAD against predB, weighted by
 * xGetMEDistortionWeight (:7666-7676).  The searched list is list 0, so the target list's weight is g_BcwWeights[bcwIdx]. */
static uint64_t vo_smvd_cost(const vo_pel* org, int orgStride, const vo_pel* refA, const vo_pel* refB, int refStride,
                             const vo_smvd_io* io, int mvAx, int mvAy, int mvBx, int mvBy)
{
  vo_pel   predA[VO_MAX_CU * VO_MAX_CU], predB[VO_MAX_CU * VO_MAX_CU], tmp[VO_MAX_CU * VO_MAX_CU];
  uint64_t dist;
  int      r;
  vo_clip_mv(&mvAx, &mvAy, io->x, io->y, io->picW, io->picH, io->maxCuW, io->maxCuH);
  vo_clip_mv(&mvBx, &mvBy, io->x, io->y, io->picW, io->picH, io->maxCuW, io->maxCuH);
  vo_mc_block(0, refA + (ptrdiff_t) io->y * refStride + io->x, refStride, io->w, io->h, mvAx, mvAy, 0, io->bd, io->imv == 3, predA, io->w);
  vo_mc_block(0, refB + (ptrdiff_t) io->y * refStride + io->x, refStride, io->w, io->h, mvBx, mvBy, 0, io->bd, io->imv == 3, predB, io->w);
  for (r = 0; r < io->h; r++) memcpy(tmp + r * io->w, org + (ptrdiff_t) r * orgStride, sizeof(vo_pel) * io->w);
  {
    static const int bcwWeights[5] = { -2, 3, 4, 5, 10 }; /* g_BcwWeights, BCW_DEFAULT = 2 */
    const int        wTar          = bcwWeights[io->bcwIdx];
    const double     fWeight       = io->bcwIdx == 2 ? 0.5 : fabs((double) wTar / 8.0);
    if (io->bcwIdx == 2)
      vo_remove_high_freq(tmp, predA, io->w * io->h, io->clipBiPred, io->bd);
    else
      vo_remove_weight_high_freq(tmp, predA, io->w * io->h, io->clipBiPred, io->bd, wTar);
    dist = io->useHad ? vo_satd(tmp, io->w, predB, io->w, io->w, io->h) : vo_sad(tmp, io->w, predB, io->w, io->w, io->h, 0);
    return (uint64_t) floor(fWeight * (double) dist);
  }
}

/* InterSearch::xSymmeticRefineMvSearch — EncoderLib/InterSearch.cpp:4393-4503 (patterns 2 = diamond and 0 = cross, the two
 * xSymmetricMotionEstimation uses; no MCTS constraint) */
static void vo_smvd_refine(const vo_pel* org, int orgStride, const vo_pel* refCur, const vo_pel* refTar, int refStride, vo_smvd_io* io,
                           int pattern, int stepShift, int maxRounds)
{
  static const int cross[4][2]   = { { 0, 1 }, { 1, 0 }, { 0, -1 }, { -1, 0 } };
  static const int diamond[8][2] = { { 0, 2 }, { 1, 1 }, { 2, 0 }, { 1, -1 }, { 0, -2 }, { -1, -1 }, { -2, 0 }, { -1, 1 } };
  const int down = io->imv == 0 ? -2 : (io->imv == 1 ? -4 : (io->imv == 2 ? -6 : -3)); /* Mv::m_amvrPrecision, Mv.cpp:43 */
  const int predX = vo_change_prec(io->curPredX, down), predY = vo_change_prec(io->curPredY, down);
  int       start = 0, end = pattern == 0 ? 3 : 7, rounding = pattern == 0 ? 4 : 8, mask = pattern == 0 ? 3 : 7;
  int       round, idx;
  for (round = 0; round < maxRounds; round++)
  {
    const int cx = io->curMvX, cy = io->curMvY; /* mvCurCenter */
    int       bestDirect = -1;
    for (idx = start; idx <= end; idx++)
    {
      const int direct = (idx + rounding) & mask;
      const int ox = (pattern == 0 ? cross[direct][0] : diamond[direct][0]) * (1 << stepShift);
      const int oy = (pattern == 0 ? cross[direct][1] : diamond[direct][1]) * (1 << stepShift);
      const int mx = cx + ox, my = cy + oy;
      const int px = io->tarPredX - (mx - io->curPredX), py = io->tarPredY - (my - io->curPredY); /* the mirrored MVD */
      uint64_t  cost = vo_mv_cost(io->lambda, vo_mv_bits(vo_change_prec(mx, down), vo_change_prec(my, down), predX, predY, 0, 0));
      cost += vo_smvd_cost(org, orgStride, refCur, refTar, refStride, io, mx, my, px, py);
      if (cost < io->cost)
      {
        io->cost   = cost;
        io->curMvX = mx;
        io->curMvY = my;
        io->tarMvX = px;
        io->tarMvY = py;
        bestDirect = direct;
      }
    }
    if (bestDirect == -1) break;
    {
      const int step = pattern == 2 ? 2 - (bestDirect & 1) : 1;
      start = bestDirect - step;
      end   = bestDirect + step;
    }
  }
}

/* InterSearch::xSymmetricMotionEstimation — EncoderLib/InterSearch.cpp:4506-4518 */
void vo_smvd_search(const vo_pel* org, int orgStride, const vo_pel* refCur, const vo_pel* refTar, int refStride, vo_smvd_io* io)
{
  const int stepShift = 2 + (io->imv == 3 ? 1 : (io->imv << 1)); /* MV_FRACTIONAL_BITS_DIFF + ... */
  vo_smvd_refine(org, orgStride, refCur, refTar, refStride, io, 2, stepShift, 8 >> io->imv);
  vo_smvd_refine(org, orgStride, refCur, refTar, refStride, io, 0, stepShift, 1);
}

/* ------------------------------------------------------------------------------------------------
 * Decoder-side MV refinement (DMVR): the search of one sub-block
 * ---------------------------------------------------------------------------------------------- */

/* Bilinear prediction of the (w+4) x (h+4) neighbourhood of the block the merge MV points at — xPrefetch
 * (InterPrediction.cpp:1664-1708: the reference samples are fetched at the integer part of clip(mv - 3 samples)),
 * xinitMC (:1949-1995: the fraction comes from clip(mv), the block starts 2 samples up and left) and xPredInterBlk with
 * bilinearMC (:660-768) on the 2-tap filter m_bilinearFilterPrec4 (InterpolationFilter.cpp:312-330) with the
 * biMCForDMVR rounding (:424-452, :600-612): 10-bit intermediates whatever the bit depth.  ref: sample (0,0) of the
 * reference plane (border extended).  dst stride w + 4. */
static void vo_dmvr_bilinear(const vo_pel* ref, int refStride, int x, int y, int w, int h, int mvx, int mvy, int picW, int picH,
                             int maxCuW, int maxCuH, int bd, vo_pel* dst)
{
  const int W2 = w + 4, H2 = h + 4;
  int       cx = mvx - 48, cy = mvy - 48, fx = mvx, fy = mvy, r, c;
  int       tmp[21][20];
  vo_clip_mv(&cx, &cy, x, y, picW, picH, maxCuW, maxCuH);
  vo_clip_mv(&fx, &fy, x, y, picW, picH, maxCuW, maxCuH);
  {
    const vo_pel* src  = ref + (ptrdiff_t) (y + (cy >> 4) + 1) * refStride + x + (cx >> 4) + 1;
    const int     xf   = fx & 15, yf = fy & 15;
    const int     sh1  = 4 - (10 - bd), off1 = 1 << (sh1 - 1); /* first pass: IF_FILTER_PREC_BILINEAR - (10 - bd) */
    if (yf == 0 && xf == 0) /* filterCopy, isFirst && !isLast */
    {
      for (r = 0; r < H2; r++)
        for (c = 0; c < W2; c++) dst[r * W2 + c] = (vo_pel) (src[(ptrdiff_t) r * refStride + c] << (10 - bd));
    }
    else if (yf == 0)
    {
      for (r = 0; r < H2; r++)
        for (c = 0; c < W2; c++)
          dst[r * W2 + c] = (vo_pel) ((src[(ptrdiff_t) r * refStride + c] * (16 - xf) + src[(ptrdiff_t) r * refStride + c + 1] * xf + off1) >> sh1);
    }
    else if (xf == 0)
    {
      for (r = 0; r < H2; r++)
        for (c = 0; c < W2; c++)
          dst[r * W2 + c] = (vo_pel) ((src[(ptrdiff_t) r * refStride + c] * (16 - yf) + src[(ptrdiff_t) (r + 1) * refStride + c] * yf + off1) >> sh1);
    }
    else
    {
      for (r = 0; r < H2 + 1; r++)
        for (c = 0; c < W2; c++)
          tmp[r][c] = (vo_pel) ((src[(ptrdiff_t) r * refStride + c] * (16 - xf) + src[(ptrdiff_t) r * refStride + c + 1] * xf + off1) >> sh1);
      for (r = 0; r < H2; r++)
        for (c = 0; c < W2; c++) dst[r * W2 + c] = (vo_pel) ((tmp[r][c] * (16 - yf) + tmp[r + 1][c] * yf + 8) >> 4);
    }
  }
}

/* div_for_maxq7 — InterPrediction.cpp:1731-1763 */
static int vo_div_maxq7(int64_t num, int64_t den)
{
  int sign = 0, q = 0;
  if (num < 0)
  {
    sign = 1;
    num  = -num;
  }
  den <<= 3;
  if (num >= den)
  {
    num -= den;
    q++;
  }
  q <<= 1;
  den >>= 1;
  if (num >= den)
  {
    num -= den;
    q++;
  }
  q <<= 1;
  if (num >= (den >> 1)) q++;
  return sign ? -q : q;
}

/* One sub-block of xProcessDMVR (InterPrediction.cpp:2098-2154): cost at the merge MVs (even rows, xDMVRCost :1919-1927),
 * biased by 1/4; if it reaches w*h, the 25 mirrored integer offsets (xBIPMVRefine :1820-1843, first strict minimum in
 * the order of m_pSearchOffset) and the parametric sub-sample step (xDMVRSubPixelErrorSurface :1929-1947,
 * xSubPelErrorSrfc :1766-1818).  out: {mvdL0SubPu.hor, .ver (1/16 sample), minCost, notZeroCost}. */
void vo_dmvr_block(const vo_pel* ref0, const vo_pel* ref1, int refStride, int x, int y, int w, int h, int mv0x, int mv0y, int mv1x,
                   int mv1y, int picW, int picH, int maxCuW, int maxCuH, int bd, int32_t* out)
{
  vo_pel    p0[20 * 20], p1[20 * 20];
  uint64_t  sad[25], minCost;
  const int W2 = w + 4;
  int       k, r, c, bestK = 12, notZero = 1, totalX = 0, totalY = 0;
  vo_dmvr_bilinear(ref0, refStride, x, y, w, h, mv0x, mv0y, picW, picH, maxCuW, maxCuH, bd, p0);
  vo_dmvr_bilinear(ref1, refStride, x, y, w, h, mv1x, mv1y, picW, picH, maxCuW, maxCuH, bd, p1);
  for (k = 0; k < 25; k++)
  {
    /* offset (ox, oy) on list 0, mirrored on list 1; rows 0, 2, 4 ... (subShift 1, the << 1 of the SAD undone by xDMVRCost) */
    const int ox = k % 5 - 2, oy = k / 5 - 2;
    uint64_t  s = 0;
    for (r = 0; r < h; r += 2)
      for (c = 0; c < w; c++) s += (uint64_t) abs(p0[(2 + oy + r) * W2 + 2 + ox + c] - p1[(2 - oy + r) * W2 + 2 - ox + c]);
    sad[k] = s;
  }
  minCost = sad[12] - (sad[12] >> 2);
  if (minCost < (uint64_t) (w * h))
    notZero = 0;
  else
  {
    sad[12] = minCost;
    if (!minCost)
      notZero = 0;
    else
      for (k = 0; k < 25; k++)
        if (sad[k] < minCost)
        {
          minCost = sad[k];
          bestK   = k;
        }
  }
  totalX = (bestK % 5 - 2) * 16;
  totalY = (bestK / 5 - 2) * 16;
  if (notZero && abs(totalX) != 32 && abs(totalY) != 32)
  {
    /* centre, left, top, right, bottom */
    const uint64_t e0 = sad[bestK], e1 = sad[bestK - 1], e2 = sad[bestK - 5], e3 = sad[bestK + 1], e4 = sad[bestK + 5];
    int64_t        num, den;
    num = (int64_t) ((e1 - e3) << 4);
    den = (int64_t) (e1 + e3 - (e0 << 1));
    if (den != 0) totalX += (e1 != e0 && e3 != e0) ? vo_div_maxq7(num, den) : (e1 == e0 ? -8 : 8);
    num = (int64_t) ((e2 - e4) << 4);
    den = (int64_t) (e2 + e4 - (e0 << 1));
    if (den != 0) totalY += (e2 != e0 && e4 != e0) ? vo_div_maxq7(num, den) : (e2 == e0 ? -8 : 8);
  }
  out[0] = totalX;
  out[1] = totalY;
  out[2] = (int32_t) minCost;
  out[3] = notZero;
}

/* The luma prediction of one list after DMVR — xFinalPaddedMCForDMVR (InterPrediction.cpp:1845-1917) over the buffer xPrefetch
 * filled (:1664-1708: the (w+7) x (h+7) window at the integer part of clip(mergeMv - 3 samples)) and xPad extended by two
 * replicated samples on every side (:1710-1730, paddingCore Buffer.cpp:340-364): the 8-tap filter of xPredInterBlk (bi = true,
 * 14-bit intermediates) at the fraction of clip(refinedMv), reading the window with clamped coordinates instead of the
 * picture.  ref: sample (0,0) of the reference plane; dst stride w. */
void vo_dmvr_final_luma(const vo_pel* ref, int refStride, int x, int y, int w, int h, int mergeX, int mergeY, int refinedX,
                        int refinedY, int picW, int picH, int maxCuW, int maxCuH, int bd, vo_pel* dst)
{
  vo_pel    patch[(16 + 7) * (16 + 7)];
  const int pw = w + 7, ph = h + 7;
  int       cx = mergeX - 48, cy = mergeY - 48, fx = refinedX, fy = refinedY, r, c;
  vo_clip_mv(&cx, &cy, x, y, picW, picH, maxCuW, maxCuH); /* xPrefetch */
  vo_clip_mv(&fx, &fy, x, y, picW, picH, maxCuW, maxCuH); /* cMvClipped: only its fraction is used */
  {
    const vo_pel* win = ref + (ptrdiff_t) (y + (cy >> 4)) * refStride + x + (cx >> 4); /* window origin */
    const int     bx = 3 + ((refinedX >> 4) - (mergeX >> 4)), by = 3 + ((refinedY >> 4) - (mergeY >> 4)); /* block origin in the window */
    for (r = 0; r < ph; r++)
      for (c = 0; c < pw; c++)
      {
        int wr = by - 3 + r, wc = bx - 3 + c; /* window coordinates of the filter's support */
        wr = wr < 0 ? 0 : (wr > h + 6 ? h + 6 : wr);
        wc = wc < 0 ? 0 : (wc > w + 6 ? w + 6 : wc);
        patch[r * pw + c] = win[(ptrdiff_t) wr * refStride + wc];
      }
  }
  vo_mc_block(0, patch + 3 * pw + 3, pw, w, h, fx & 15, fy & 15, 1, bd, 0, dst, w);
}

/* The same for a 4:2:0 chroma component when the block moved (an unmoved block reads the picture directly, which is the
 * plain xPredInterBlk): window of (w/2+3) x (h/2+3) chroma samples at the integer part (1/32 units) of clip(mergeMv - 1
 * chroma sample), padded by one sample, 4-tap filter.  x, y, w, h: the LUMA rectangle; ref: sample (0,0) of the chroma plane. */
void vo_dmvr_final_chroma(const vo_pel* ref, int refStride, int x, int y, int w, int h, int mergeX, int mergeY, int refinedX,
                          int refinedY, int picW, int picH, int maxCuW, int maxCuH, int bd, vo_pel* dst)
{
  vo_pel    patch[(8 + 3) * (8 + 3)];
  const int cw = w >> 1, ch = h >> 1, pw = cw + 3, ph = ch + 3;
  int       cx = mergeX - 32, cy = mergeY - 32, fx = refinedX, fy = refinedY, r, c;
  vo_clip_mv(&cx, &cy, x, y, picW, picH, maxCuW, maxCuH);
  vo_clip_mv(&fx, &fy, x, y, picW, picH, maxCuW, maxCuH);
  {
    const vo_pel* win = ref + (ptrdiff_t) ((y >> 1) + (cy >> 5)) * refStride + (x >> 1) + (cx >> 5);
    const int     bx = 1 + ((refinedX >> 5) - (mergeX >> 5)), by = 1 + ((refinedY >> 5) - (mergeY >> 5));
    for (r = 0; r < ph; r++)
      for (c = 0; c < pw; c++)
      {
        int wr = by - 1 + r, wc = bx - 1 + c;
        wr = wr < 0 ? 0 : (wr > ch + 2 ? ch + 2 : wr);
        wc = wc < 0 ? 0 : (wc > cw + 2 ? cw + 2 : wc);
        patch[r * pw + c] = win[(ptrdiff_t) wr * refStride + wc];
      }
  }
  vo_mc_block(1, patch + pw + 1, pw, cw, ch, fx & 31, fy & 31, 1, bd, 0, dst, cw);
}

/* ------------------------------------------------------------------------------------------------
 * GOP-based temporal filter: motion estimation
 * ---------------------------------------------------------------------------------------------- */

#define VO_MCTF_PAD 128 /* EncTemporalFilter::m_padding */

static const int vo_mctf_filter[16][8] = { /* EncTemporalFilter::m_interpolationFilter — EncTemporalFilter.cpp:50-68 */
  { 0, 0, 0, 64, 0, 0, 0, 0 },    { 0, 1, -3, 64, 4, -2, 0, 0 },   { 0, 1, -6, 62, 9, -3, 1, 0 },   { 0, 2, -8, 60, 14, -5, 1, 0 },
  { 0, 2, -9, 57, 19, -7, 2, 0 }, { 0, 3, -10, 53, 24, -8, 2, 0 }, { 0, 3, -11, 50, 29, -9, 2, 0 }, { 0, 3, -11, 44, 35, -10, 3, 0 },
  { 0, 1, -7, 38, 38, -7, 1, 0 }, { 0, 3, -10, 35, 44, -11, 3, 0 }, { 0, 2, -9, 29, 50, -11, 3, 0 }, { 0, 2, -8, 24, 53, -10, 3, 0 },
  { 0, 2, -7, 19, 57, -9, 2, 0 }, { 0, 1, -5, 14, 60, -8, 2, 0 },  { 0, 1, -3, 9, 62, -6, 1, 0 },   { 0, 0, -2, 4, 64, -3, 1, 0 }
};

/* EncTemporalFilter::motionErrorLuma — EncTemporalFilter.cpp:268-361 */
int vo_mctf_error(const vo_pel* org, int orgStride, const vo_pel* ref, int refStride, int x, int y, int dx, int dy, int bs,
                  int bestError, int bitDepth)
{
  int error = 0, x1, y1;
  if (((dx | dy) & 0xF) == 0)
  {
    dx /= 16;
    dy /= 16;
    for (y1 = 0; y1 < bs; y1++)
    {
      const vo_pel* o = org + (ptrdiff_t) (y + y1) * orgStride + x;
      const vo_pel* b = ref + (ptrdiff_t) (y + y1 + dy) * refStride + (x + dx);
      for (x1 = 0; x1 < bs; x1++)
      {
        const int diff = o[x1] - b[x1];
        error += diff * diff;
      }
      if (error > bestError) return error;
    }
  }
  else
  {
    const int* xf = vo_mctf_filter[dx & 0xF];
    const int* yf = vo_mctf_filter[dy & 0xF];
    const int  maxv = (1 << bitDepth) - 1;
    int        tmp[64 + 8][64];
    for (y1 = 1; y1 < bs + 7; y1++)
    {
      const vo_pel* row = ref + (ptrdiff_t) (y + y1 + (dy >> 4) - 3) * refStride;
      for (x1 = 0; x1 < bs; x1++)
      {
        const vo_pel* p   = row + (x + x1 + (dx >> 4) - 3);
        int           sum = 0, k;
        for (k = 1; k <= 6; k++) sum += xf[k] * p[k];
        tmp[y1][x1] = sum;
      }
    }
    for (y1 = 0; y1 < bs; y1++)
    {
      const vo_pel* o = org + (ptrdiff_t) (y + y1) * orgStride;
      for (x1 = 0; x1 < bs; x1++)
      {
        int sum = 0, k;
        for (k = 1; k <= 6; k++) sum += yf[k] * tmp[y1 + k][x1];
        sum = (sum + (1 << 11)) >> 12;
        sum = sum < 0 ? 0 : (sum > maxv ? maxv : sum);
        error += (sum - o[x + x1]) * (sum - o[x + x1]);
      }
      if (error > bestError) return error;
    }
  }
  return error;
}

typedef struct
{
  vo_pel* base;
  vo_pel* origin;
  int     stride, width, height;
} vo_plane;

static void vo_plane_alloc(vo_plane* p, int width, int height)
{
  p->width  = width;
  p->height = height;
  p->stride = width + 2 * VO_MCTF_PAD;
  p->base   = (vo_pel*) malloc(sizeof(vo_pel) * (size_t) p->stride * (height + 2 * VO_MCTF_PAD));
  p->origin = p->base + (size_t) VO_MCTF_PAD * p->stride + VO_MCTF_PAD;
}

/* EncTemporalFilter::subsampleLuma — EncTemporalFilter.cpp:241-266 (2x2 mean, then extendBorderPel) */
static void vo_mctf_subsample(const vo_pel* in, int inStride, int inW, int inH, vo_plane* out)
{
  int x, y;
  vo_plane_alloc(out, inW / 2, inH / 2);
  for (y = 0; y < out->height; y++)
    for (x = 0; x < out->width; x++)
    {
      const vo_pel* a = in + (ptrdiff_t) (2 * y) * inStride + 2 * x;
      out->origin[(ptrdiff_t) y * out->stride + x] = (vo_pel) ((a[0] + a[inStride] + a[1] + a[inStride + 1] + 2) >> 2);
    }
  for (y = -VO_MCTF_PAD; y < out->height + VO_MCTF_PAD; y++)
    for (x = -VO_MCTF_PAD; x < out->width + VO_MCTF_PAD; x++)
    {
      const int sx = x < 0 ? 0 : (x >= out->width ? out->width - 1 : x);
      const int sy = y < 0 ? 0 : (y >= out->height ? out->height - 1 : y);
      out->origin[(ptrdiff_t) y * out->stride + x] = out->origin[(ptrdiff_t) sy * out->stride + sx];
    }
}

typedef struct
{
  int x, y, error;
} vo_mctf_mv;

/* EncTemporalFilter::motionEstimationLuma — EncTemporalFilter.cpp:363-446.  mvs / previous: arrays of mvW entries per row */
static void vo_mctf_level(vo_mctf_mv* mvs, int mvW, const vo_pel* org, int os, const vo_pel* buf, int bstride, int origWidth,
                          int origHeight, int blockSize, const vo_mctf_mv* previous, int prevW, int factor, int doubleRes, int bd)
{
  int range = 5, blockX, blockY, px, py, x2, y2;
  for (blockY = 0; blockY + blockSize < origHeight; blockY += blockSize)
    for (blockX = 0; blockX + blockSize < origWidth; blockX += blockSize)
    {
      vo_mctf_mv best = { 0, 0, INT32_MAX }, prev;
      if (!previous)
        range = 8;
      else
        for (py = -2; py <= 2; py++)
        {
          const int testy = blockY / (2 * blockSize) + py;
          for (px = -2; px <= 2; px++)
          {
            const int testx = blockX / (2 * blockSize) + px;
            if (testx >= 0 && testx < origWidth / (2 * blockSize) && testy >= 0 && testy < origHeight / (2 * blockSize))
            {
              const vo_mctf_mv old = previous[testy * prevW + testx];
              const int        e   = vo_mctf_error(org, os, buf, bstride, blockX, blockY, old.x * factor, old.y * factor, blockSize, best.error, bd);
              if (e < best.error)
              {
                best.x     = old.x * factor;
                best.y     = old.y * factor;
                best.error = e;
              }
            }
          }
        }
      prev = best;
      for (y2 = prev.y / 16 - range; y2 <= prev.y / 16 + range; y2++)
        for (x2 = prev.x / 16 - range; x2 <= prev.x / 16 + range; x2++)
        {
          const int e = vo_mctf_error(org, os, buf, bstride, blockX, blockY, x2 * 16, y2 * 16, blockSize, best.error, bd);
          if (e < best.error)
          {
            best.x     = x2 * 16;
            best.y     = y2 * 16;
            best.error = e;
          }
        }
      if (doubleRes)
      {
        int pass;
        for (pass = 0; pass < 2; pass++) /* +-12 in steps of 4, then +-3 in steps of 1 (1/16 sample) */
        {
          const int dr = pass ? 3 : 12, st = pass ? 1 : 4;
          prev = best;
          for (y2 = prev.y - dr; y2 <= prev.y + dr; y2 += st)
            for (x2 = prev.x - dr; x2 <= prev.x + dr; x2 += st)
            {
              const int e = vo_mctf_error(org, os, buf, bstride, blockX, blockY, x2, y2, blockSize, best.error, bd);
              if (e < best.error)
              {
                best.x     = x2;
                best.y     = y2;
                best.error = e;
              }
            }
        }
      }
      mvs[(blockY / blockSize) * mvW + blockX / blockSize] = best;
    }
}

/* EncTemporalFilter::motionEstimation — EncTemporalFilter.cpp:448-466 */
void vo_mctf_me(const vo_pel* org, int orgStride, const vo_pel* ref, int refStride, int width, int height, int bitDepth,
                int32_t* mvOut)
{
  const int   lw = width / 16, lh = height / 16; /* the three intermediate fields are (width/16) x (height/16) */
  const int   fw = width / 4, fh = height / 4;
  vo_plane    o2, o4, b2, b4;
  vo_mctf_mv *mv0, *mv1, *mv2, *mv;
  int         i;
  mv0 = (vo_mctf_mv*) malloc(sizeof(vo_mctf_mv) * (size_t) lw * lh);
  mv1 = (vo_mctf_mv*) malloc(sizeof(vo_mctf_mv) * (size_t) lw * lh);
  mv2 = (vo_mctf_mv*) malloc(sizeof(vo_mctf_mv) * (size_t) lw * lh);
  mv  = (vo_mctf_mv*) malloc(sizeof(vo_mctf_mv) * (size_t) fw * fh);
  for (i = 0; i < lw * lh; i++)
  {
    const vo_mctf_mv d = { 0, 0, INT32_MAX };
    mv0[i] = mv1[i] = mv2[i] = d;
  }
  for (i = 0; i < fw * fh; i++)
  {
    const vo_mctf_mv d = { 0, 0, INT32_MAX };
    mv[i] = d;
  }
  vo_mctf_subsample(org, orgStride, width, height, &o2);
  vo_mctf_subsample(o2.origin, o2.stride, o2.width, o2.height, &o4);
  vo_mctf_subsample(ref, refStride, width, height, &b2);
  vo_mctf_subsample(b2.origin, b2.stride, b2.width, b2.height, &b4);
  vo_mctf_level(mv0, lw, o4.origin, o4.stride, b4.origin, b4.stride, o4.width, o4.height, 16, NULL, 0, 1, 0, bitDepth);
  vo_mctf_level(mv1, lw, o2.origin, o2.stride, b2.origin, b2.stride, o2.width, o2.height, 16, mv0, lw, 2, 0, bitDepth);
  vo_mctf_level(mv2, lw, org, orgStride, ref, refStride, width, height, 16, mv1, lw, 2, 0, bitDepth);
  vo_mctf_level(mv, fw, org, orgStride, ref, refStride, width, height, 8, mv2, lw, 1, 1, bitDepth);
  for (i = 0; i < fw * fh; i++)
  {
    mvOut[3 * i]     = mv[i].x;
    mvOut[3 * i + 1] = mv[i].y;
    mvOut[3 * i + 2] = mv[i].error;
  }
  free(mv0);
  free(mv1);
  free(mv2);
  free(mv);
  free(o2.base);
  free(o4.base);
  free(b2.base);
  free(b4.base);
}

/* EncTemporalFilter::applyMotion — EncTemporalFilter.cpp:470-552, one component */
void vo_mctf_apply_motion(const vo_pel* src, int srcStride, int compW, int compH, int csx, int csy, const int32_t* mv,
                          int mvStride, int bitDepth, vo_pel* dst, int dstStride)
{
  const int bsx = 8 >> csx, bsy = 8 >> csy, maxv = (1 << bitDepth) - 1;
  int       x, y, bnx, bny, bx, by, k;
  for (y = 0, bny = 0; y + bsy <= compH; y += bsy, bny++)
    for (x = 0, bnx = 0; x + bsx <= compW; x += bsx, bnx++)
    {
      const int32_t* m    = mv + 3 * ((size_t) bny * mvStride + bnx);
      const int      dx = m[0] >> csx, dy = m[1] >> csy;
      const int      xInt = m[0] >> (4 + csx), yInt = m[1] >> (4 + csy);
      const int*     xf = vo_mctf_filter[dx & 0xf];
      const int*     yf = vo_mctf_filter[dy & 0xf];
      int            tmp[8 + 7][8];
      for (by = 1; by < bsy + 7; by++)
      {
        const vo_pel* row = src + (ptrdiff_t) (y + by + yInt - 3) * srcStride;
        for (bx = 0; bx < bsx; bx++)
        {
          const vo_pel* p   = row + (x + bx + xInt - 3);
          int           sum = 0;
          for (k = 1; k <= 6; k++) sum += xf[k] * p[k];
          tmp[by][bx] = sum;
        }
      }
      for (by = 0; by < bsy; by++)
        for (bx = 0; bx < bsx; bx++)
        {
          int sum = 0;
          for (k = 1; k <= 6; k++) sum += yf[k] * tmp[by + k][bx];
          sum = (sum + (1 << 11)) >> 12;
          dst[(ptrdiff_t) (y + by) * dstStride + x + bx] = (vo_pel) (sum < 0 ? 0 : (sum > maxv ? maxv : sum));
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Affine motion estimation: the three dispatch-table primitives of AffineGradientSearch
 * (CommonLib/AffineGradientSearch.cpp:64-174; SIMD x86/AffineGradientSearchX86.h), called once per iteration of
 * InterSearch::xAffineMotionEstimation (EncoderLib/InterSearch.cpp:5501-5526)
 * ---------------------------------------------------------------------------------------------- */

/* xHorizontalSobelFilter / xVerticalSobelFilter: 3x3 Sobel of the prediction; the border takes the nearest interior value */
void vo_affine_sobel(int vertical, const vo_pel* pred, int predStride, int* deriv, int derivStride, int w, int h)
{
  int j, k;
  for (j = 0; j < h; j++)
    for (k = 0; k < w; k++)
    {
      const int jj = j < 1 ? 1 : (j > h - 2 ? h - 2 : j), kk = k < 1 ? 1 : (k > w - 2 ? w - 2 : k);
      const vo_pel* c = pred + (ptrdiff_t) jj * predStride + kk;
      int v;
      if (!vertical)
        v = c[1 - predStride] - c[-1 - predStride] + (c[1] << 1) - (c[-1] << 1) + c[1 + predStride] - c[-1 + predStride];
      else
        v = c[predStride - 1] - c[-predStride - 1] + (c[predStride] << 1) - (c[-predStride] << 1) + c[predStride + 1] - c[-predStride + 1];
      deriv[(ptrdiff_t) j * derivStride + k] = v;
    }
}

/* xEqualCoeffComputer: accumulates into coeff[7][7] (rows 1..n, columns 0..n; n = 4 or 6 affine parameters) */
void vo_affine_equal_coeff(const vo_pel* residue, int residueStride, const int* d0, const int* d1, int derivStride, int64_t coeff[7][7],
                           int w, int h, int sixParam)
{
  const int n = sixParam ? 6 : 4;
  int j, k, col, row;
  for (j = 0; j < h; j++)
  {
    const int cy = ((j >> 2) << 2) + 2;
    for (k = 0; k < w; k++)
    {
      const int cx = ((k >> 2) << 2) + 2, gx = d0[(ptrdiff_t) j * derivStride + k], gy = d1[(ptrdiff_t) j * derivStride + k];
      int c[6];
      if (!sixParam)
      {
        c[0] = gx;
        c[1] = cx * gx + cy * gy;
        c[2] = gy;
        c[3] = cy * gx - cx * gy;
      }
      else
      {
        c[0] = gx;
        c[1] = cx * gx;
        c[2] = gy;
        c[3] = cx * gy;
        c[4] = cy * gx;
        c[5] = cy * gy;
      }
      for (col = 0; col < n; col++)
      {
        for (row = 0; row < n; row++) coeff[col + 1][row] += (int64_t) c[col] * c[row];
        coeff[col + 1][n] += ((int64_t) c[col] * residue[(ptrdiff_t) j * residueStride + k]) << 3;
      }
    }
  }
}
