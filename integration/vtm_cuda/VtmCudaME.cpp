// See VtmCudaME.h.  Host side of the drop-in, written in the reference's language (C++11) over the C ABI
// (include/vtmme.h).
#include "VtmCudaME.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "CommonLib/CommonDef.h"
#include "CommonLib/Picture.h"
#include "vtmme.h"

namespace vtmcuda
{
namespace
{
vtmme_ctx* g_ctx        = nullptr;
uint64_t   g_calls      = 0;
uint64_t   g_uploads    = 0;
uint64_t   g_tzSearches = 0;   // searches whose integer stage was xTZSearch / xTZSearchSelective (FastSearch=1/2/3)
uint64_t   g_intRefines = 0;   // searches whose xPatternSearchIntRefine ran on the GPU too
double     g_searchSec  = 0;   // wall time spent inside vtmme_search (upload of the pattern, kernels, sync)
int        g_nextPicId  = 1;
uint64_t   g_batchCalls = 0;   // vtmme_search calls that carried the collected searches of a PU
uint64_t   g_batchJobs  = 0;   // searches run in such calls
uint64_t   g_batchHits  = 0;   // requests answered from a batch
bool       g_collecting = false;
uint64_t   g_tfPics     = 0;   // pictures filtered by the GOP-based temporal filter on the GPU
uint64_t   g_tfRefs     = 0;   // neighbouring pictures searched for them
double     g_tfSec      = 0;

struct Prefetched
{
  alignas( 8 ) unsigned char raw[sizeof( vtmcuda::SearchIn )];   // the request byte for byte (the caller zero-fills it first)
  vtmcuda::SearchOut         out;
  bool                       valid;
  const vtmcuda::SearchIn&   in() const { return *reinterpret_cast<const vtmcuda::SearchIn*>( raw ); }
};
std::vector<Prefetched> g_prefetched;

struct Uploaded
{
  int id;
  int poc;
};
std::map<const Picture*, Uploaded> g_pics;   // Picture objects are recycled for new POCs: keyed by (object, POC)

int envFlag( const char* name )
{
  const char* e = getenv( name );
  return e && e[0] && e[0] != '0';
}

vtmme_ctx* ctx()
{
  if( !g_ctx )
  {
    const char* d  = getenv( "VTMME_DEVICE" );
    const int   rc = vtmme_create( d ? atoi( d ) : 0, &g_ctx );
    CHECK( rc != VTMME_OK, "vtmme_create failed (no CUDA device?) — the GPU motion search has no CPU fallback" );
    atexit( printStats );
  }
  return g_ctx;
}

// device copy of a reference picture's luma recon plane; uploaded once per (picture, POC)
int pictureId( const Picture* pic )
{
  auto it = g_pics.find( pic );
  if( it != g_pics.end() && it->second.poc == pic->getPOC() )
  {
    return it->second.id;
  }
  const int     id  = it != g_pics.end() ? it->second.id : g_nextPicId++;
  const CPelBuf buf = pic->getRecoBuf( COMPONENT_Y );   // (0,0) of the picture; border extended by extendPicBorder
  const int     rc  = vtmme_upload_picture( ctx(), id, buf.buf, buf.stride, buf.width, buf.height, (int) pic->margin, 1 );
  CHECK( rc != VTMME_OK, vtmme_last_error( g_ctx ) );
  g_pics[pic] = Uploaded{ id, pic->getPOC() };
  g_uploads++;
  return id;
}
}   // namespace

bool enabled()
{
  static int e = -1;
  if( e < 0 ) e = envFlag( "VTMME_ENABLE" );
  return e != 0;
}

bool tableHooksEnabled()
{
  static int e = -1;
  if( e < 0 ) e = envFlag( "VTMME_TABLE_HOOKS" );
  return e != 0;
}

namespace
{
// SearchIn -> vtmme_job (+ the AMVR / TZ descriptors it points at)
void fillJob( const SearchIn& in, vtmme_job& j, vtmme_amvr& a, vtmme_tz& t )
{
  j.curPic   = 0;
  j.refPic   = pictureId( in.refPic );
  j.x        = in.x;
  j.y        = in.y;
  j.w        = in.w;
  j.h        = in.h;
  j.org      = in.org;
  j.orgStride = in.orgStride;
  j.srLeft   = in.srLeft;
  j.srRight  = in.srRight;
  j.srTop    = in.srTop;
  j.srBottom = in.srBottom;
  j.predQx   = in.predQx;
  j.predQy   = in.predQy;
  j.imvShift = in.imvShift;
  // DistParam::subShift as RdCost::setDistParam derives it (RdCost.cpp:289-323); the full search uses mode 0 or 2, the
  // TZ searches of MESEARCH_SELECTIVE mode 1 (staged SAD, InterSearch.cpp:3438-3445)
  j.subShift = 0;
  if( in.subShiftMode == 2 && in.h > 8 && in.w <= 64 ) j.subShift = 1;
  if( in.subShiftMode == 1 )
  {
    j.subShift = ( in.h > 32 && ( in.h & 15 ) == 0 ) ? 4 : ( in.h > 16 && ( in.h & 7 ) == 0 ) ? 3 : ( in.h > 8 && ( in.h & 3 ) == 0 ) ? 2 : ( ( in.h & 1 ) == 0 ? 1 : 0 );
  }
  CHECK( in.subShiftMode != 0 && in.subShiftMode != 2 && !( in.subShiftMode == 1 && in.tzSearch ), "unexpected subShiftMode" );
  j.bitDepth     = in.bitDepth;
  j.useHad       = in.useHad;
  j.useAltHpel   = in.useAltHpel;
  j.fracMode     = in.doFrac;
  j.lambdaMotion = in.lambdaMotion;
  j.amvr = nullptr;
  if( in.doFrac == 2 )
  {
    a.imv     = in.imv;
    a.numCand = in.numCand;
    for( int i = 0; i < 2; i++ )
    {
      a.candX[i]      = in.candX[i];
      a.candY[i]      = in.candY[i];
      a.mvpIdxBits[i] = in.mvpIdxBits[i];
    }
    a.mvpIdx  = in.mvpIdx;
    a.bits    = in.bits;
    a.picW    = in.picW;
    a.picH    = in.picH;
    a.maxCuW  = in.maxCuW;
    a.maxCuH  = in.maxCuH;
    a.fWeight = in.fWeight;
    j.amvr    = &a;
  }
  j.tz = nullptr;
  if( in.tzSearch )
  {
    t.startX      = in.tzStartX;
    t.startY      = in.tzStartY;
    t.hasInt2Nx2N = 0;   // xMotionEstimation always passes pIntegerMv2Nx2NPred = 0 (InterSearch.cpp:3453,3455)
    t.int2Nx2NX = t.int2Nx2NY = 0;
    t.nSeeds      = in.tzNumSeeds;
    for( int i = 0; i < 16; i++ )
    {
      t.seedX[i] = i < in.tzNumSeeds ? in.tzSeedX[i] : 0;
      t.seedY[i] = i < in.tzNumSeeds ? in.tzSeedY[i] : 0;
    }
    t.searchRange     = in.tzSearchRange;
    t.extended        = in.tzExtended;
    t.fast            = in.tzFast;
    t.firstSearchStop = in.tzFirstSearchStop;
    t.picW            = in.picW;
    t.picH            = in.picH;
    t.maxCu           = in.maxCuW;
    t.selective       = in.tzSelective;
    t.stagedSad       = in.subShiftMode == 1;
    j.tz              = &t;
  }
}

void fillOut( const vtmme_result& r, SearchOut& out )
{
  out.mvX      = r.mvX;
  out.mvY      = r.mvY;
  out.intSad   = r.intSad;
  out.halfX    = r.halfX;
  out.halfY    = r.halfY;
  out.qterX    = r.qterX;
  out.qterY    = r.qterY;
  out.fracCost = r.fracCost;
  out.amvrMvX  = r.amvrMvX;
  out.amvrMvY  = r.amvrMvY;
  out.mvpIdx   = r.mvpIdx;
  out.bits     = r.bits;
  out.cost     = r.cost;
}
}   // namespace

bool batching()
{
  static int e = -1;
  if( e < 0 )
  {
    const char* v = getenv( "VTMME_BATCH" );
    e             = ( v && v[0] == '0' ) ? 0 : 1;
  }
  return e != 0;
}

bool collecting() { return g_collecting; }

void beginCollect()
{
  g_prefetched.clear();
  g_collecting = true;
}

void endCollect()
{
  g_collecting = false;
  // the library takes full-search and TZ jobs in separate calls
  for( int kind = 0; kind < 2; kind++ )
  {
    std::vector<int> idx;
    for( size_t i = 0; i < g_prefetched.size(); i++ )
      if( (int) g_prefetched[i].in().tzSearch == kind ) idx.push_back( (int) i );
    if( idx.size() < 2 ) continue;   // a single search gains nothing from being early
    const int                 n = (int) idx.size();
    std::vector<vtmme_job>    jobs( n );
    std::vector<vtmme_amvr>   amvr( n );
    std::vector<vtmme_tz>     tz( n );
    std::vector<vtmme_result> res( n );
    for( int k = 0; k < n; k++ ) fillJob( g_prefetched[idx[k]].in(), jobs[k], amvr[k], tz[k] );
    vtmme_ctx* c  = ctx();
    const auto t0 = std::chrono::steady_clock::now();
    const int  rc = vtmme_search( c, jobs.data(), n, res.data() );
    g_searchSec += std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
    CHECK( rc != VTMME_OK, vtmme_last_error( g_ctx ) );
    for( int k = 0; k < n; k++ )
    {
      fillOut( res[k], g_prefetched[idx[k]].out );
      g_prefetched[idx[k]].valid = true;
    }
    g_batchCalls++;
    g_batchJobs += n;
  }
}

void search( const SearchIn& in, SearchOut& out )
{
  if( g_collecting )
  {
    Prefetched p;
    memcpy( p.raw, &in, sizeof( SearchIn ) );
    p.valid = false;
    memset( &p.out, 0, sizeof( p.out ) );
    out = p.out;
    g_prefetched.push_back( p );
    return;
  }
  for( size_t i = 0; i < g_prefetched.size(); i++ )
  {
    // the caller zero-fills SearchIn before setting its fields: identical inputs are identical bytes
    if( g_prefetched[i].valid && memcmp( g_prefetched[i].raw, &in, sizeof( SearchIn ) ) == 0 )
    {
      out                   = g_prefetched[i].out;
      g_prefetched[i].valid = false;
      g_batchHits++;
      g_calls++;
      if( in.tzSearch ) g_tzSearches++;
      if( in.doFrac == 2 ) g_intRefines++;
      return;
    }
  }
  vtmme_job  j;
  vtmme_amvr a;
  vtmme_tz   t;
  fillJob( in, j, a, t );
  if( in.tzSearch ) g_tzSearches++;
  vtmme_result r;
  vtmme_ctx*   c  = ctx();
  const auto   t0 = std::chrono::steady_clock::now();
  const int    rc = vtmme_search( c, &j, 1, &r );
  g_searchSec += std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
  CHECK( rc != VTMME_OK, vtmme_last_error( g_ctx ) );
  fillOut( r, out );
  g_calls++;
  if( in.doFrac == 2 ) g_intRefines++;
}

void temporalFilter( int numRefs, int numComp, const TfPlane* org, const TfPlane* refs, int csx, int csy, const int* bitDepth,
                     const double* const* weights, int16_t* const* dst, const int* dstStride )
{
  vtmme_ctx* c = ctx();
  const auto t0 = std::chrono::steady_clock::now();
  // picture ids of the temporal filter live above the reference-picture cache's
  const int idOrg = 1 << 20, idRef = idOrg + 16, idCorr = idRef + 256;
  CHECK( numRefs < 1 || numRefs > 8 || numComp < 1 || numComp > 3, "temporal filter: 1..8 neighbouring pictures, 1..3 components" );
  for( int k = 0; k < numComp; k++ )
  {
    CHECK( vtmme_upload_picture( c, idOrg + k, org[k].buf, org[k].stride, org[k].width, org[k].height, 0, 0 ) != VTMME_OK, vtmme_last_error( c ) );
    for( int r = 0; r < numRefs; r++ )
    {
      const TfPlane& p = refs[r * numComp + k];
      CHECK( vtmme_upload_picture( c, idRef + r * 4 + k, p.buf, p.stride, p.width, p.height, 0, 0 ) != VTMME_OK, vtmme_last_error( c ) );
    }
  }
  // EncTemporalFilter::motionEstimation of every neighbour (:207), one call
  const int            w = org[0].width, h = org[0].height, mvW = w / 4, mvH = h / 4;
  std::vector<int32_t> orgIds( numRefs, idOrg ), refIds( numRefs ), mv( (size_t) numRefs * mvW * mvH * 3 );
  for( int r = 0; r < numRefs; r++ ) refIds[r] = idRef + r * 4;
  CHECK( vtmme_mctf_me( c, numRefs, orgIds.data(), refIds.data(), bitDepth[0], mv.data() ) != VTMME_OK, vtmme_last_error( c ) );
  // applyMotion (:564-567) and the weighting (:580-621) per component
  for( int k = 0; k < numComp; k++ )
  {
    const int            cw = org[k].width, ch = org[k].height, sx = k ? csx : 0, sy = k ? csy : 0;
    std::vector<int16_t> corr( (size_t) cw * ch ), out( (size_t) cw * ch );
    std::vector<int32_t> corrIds( numRefs );
    for( int r = 0; r < numRefs; r++ )
    {
      CHECK( vtmme_mctf_apply_motion( c, idRef + r * 4 + k, sx, sy, mv.data() + (size_t) r * mvW * mvH * 3, mvW, mvH, bitDepth[k], corr.data() ) != VTMME_OK,
             vtmme_last_error( c ) );
      corrIds[r] = idCorr + r * 4 + k;
      CHECK( vtmme_upload_picture( c, corrIds[r], corr.data(), cw, cw, ch, 0, 0 ) != VTMME_OK, vtmme_last_error( c ) );
    }
    CHECK( vtmme_mctf_bilateral( c, idOrg + k, numRefs, corrIds.data(), weights[k], bitDepth[k], out.data() ) != VTMME_OK, vtmme_last_error( c ) );
    for( int y = 0; y < ch; y++ ) memcpy( dst[k] + (ptrdiff_t) y * dstStride[k], out.data() + (size_t) y * cw, (size_t) cw * 2 );
  }
  g_tfSec += std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
  g_tfPics++;
  g_tfRefs += numRefs;
}

uint64_t distHost( int kind, const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w, int h, int subShift )
{
  uint64_t  v  = 0;
  const int rc = vtmme_dist_host( ctx(), kind, org, orgStride, cur, curStride, w, h, subShift, &v );
  CHECK( rc != VTMME_OK, vtmme_last_error( g_ctx ) );
  return v;
}

void interpHost( int comp, int vertical, const int16_t* src, int srcStride, int16_t* dst, int dstStride, int w, int h,
                 int frac, int isFirst, int isLast, int bitDepth, int useAltHpel )
{
  const int rc = vtmme_interp_host( ctx(), comp, vertical, src, srcStride, dst, dstStride, w, h, frac, isFirst, isLast,
                                    bitDepth, useAltHpel );
  CHECK( rc != VTMME_OK, vtmme_last_error( g_ctx ) );
}

void filterHost( int taps, int vertical, int isFirst, int isLast, int copy, const int16_t* src, int srcStride, int16_t* dst,
                 int dstStride, int w, int h, const int16_t* coeff, int bitDepth )
{
  const int rc = vtmme_filter_host( ctx(), taps, vertical, isFirst, isLast, copy, src, srcStride, dst, dstStride, w, h, coeff,
                                    bitDepth );
  CHECK( rc != VTMME_OK, vtmme_last_error( g_ctx ) );
}

void affineSobelHost( int vertical, const int16_t* pred, int predStride, int w, int h, int* deriv, int derivStride )
{
  const int rc = vtmme_affine_sobel_host( ctx(), vertical, pred, predStride, w, h, deriv, derivStride );
  CHECK( rc != VTMME_OK, vtmme_last_error( g_ctx ) );
}

void affineEqualCoeffHost( const int16_t* residue, int residueStride, const int* d0, const int* d1, int derivStride, int w, int h, int sixParam,
                           int64_t* coeff )
{
  const int rc = vtmme_affine_equal_coeff_host( ctx(), residue, residueStride, d0, d1, derivStride, w, h, sixParam, coeff );
  CHECK( rc != VTMME_OK, vtmme_last_error( g_ctx ) );
}

void printStats()
{
  if( g_ctx )
  {
    const uint64_t libCalls = g_calls - g_batchHits + g_batchCalls;   // vtmme_search calls actually made
    fprintf( stderr, "[vtmcuda] GPU motion searches: %llu (%llu TZ searches, %llu with AMVR integer refinement; %.1f s inside vtmme_search, %.1f us per "
                     "search), reference pictures uploaded: %llu, kernel launches: %llu\n",
             (unsigned long long) g_calls, (unsigned long long) g_tzSearches, (unsigned long long) g_intRefines, g_searchSec, g_calls ? 1e6 * g_searchSec / g_calls : 0.0,
             (unsigned long long) g_uploads, (unsigned long long) vtmme_launch_count( g_ctx ) );
    fprintf( stderr, "[vtmcuda] per-PU batching: %llu vtmme_search calls in all, %llu of them batches carrying %llu searches, %llu searches answered "
                     "from a batch (%llu collected searches were never asked for)\n",
             (unsigned long long) libCalls, (unsigned long long) g_batchCalls, (unsigned long long) g_batchJobs, (unsigned long long) g_batchHits,
             (unsigned long long) ( g_batchJobs - g_batchHits ) );
    if( g_tfPics )
      fprintf( stderr, "[vtmcuda] temporal filter: %llu pictures filtered on the GPU against %llu neighbouring pictures, %.2f s\n",
               (unsigned long long) g_tfPics, (unsigned long long) g_tfRefs, g_tfSec );
  }
}
}   // namespace vtmcuda

