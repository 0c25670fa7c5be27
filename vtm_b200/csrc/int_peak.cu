// INT/ALU-pipe peak microbenchmark for sm_100a (the roofline denominator of the motion search
// is the measured issue rate of the SAD instruction, not a datasheet number).
//
// Each kernel runs a long dependency-free stream of one instruction class (8-16 independent accumulators
// per thread, 256 threads/CTA, 8 CTAs/SM resident) and reports lane-instructions / clk / SM from the
// elapsed time and the SM clock it ran at (clock64 inside the kernel).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "vtmme_internal.h"

namespace {

constexpr int kUnroll = 16;   // independent accumulators per thread

// Variant ids (keep in sync with vtm_b200/peaks.py)
enum { V_VABSDIFF = 0, V_IADD3, V_IMAD, V_LOP3, V_PRMT, V_VABSDIFF_IMAD, V_FADD_ABS, V_VABSDIFF_FADD, V_VIADD16X2,
       V_VIADDMNMX16X2, V_VABSDIFF4, V_VABSDIFF_FADD2, V_DP2A, V_DP4A, V_FADD2P, V_VABSDIFF_FADD2P_88, V_VABSDIFF_FADD2P_610,
       V_VABSDIFF_FADD2P_106, V_VABSDIFF_FADD2P_124, V_MIX_FADD_124, V_MIX_FADD_106, V_MIX_FADD_88, V_COUNT };

// Packed FP32 (sm_100 add.f32x2 / sub.f32x2 -> FADD2): two |a-b| per instruction pair.  ptxas folds the broadcast of the
// scalar minuend, the negation and the magnitude into operand modifiers (FADD2 R, Ra.F32, -Rb.F32x2 ; FADD2 R, Racc, |Rd|).
__device__ __forceinline__ unsigned long long f2_pack(float a, float b)
{
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_absdiff_acc(unsigned long long acc, float o, unsigned long long p2)
{
  unsigned long long d;
  const unsigned long long oo = f2_pack(o, o);
  asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(oo), "l"(p2));
  float d0, d1;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
  const unsigned long long ad = f2_pack(fabsf(d0), fabsf(d1));
  asm("add.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(ad));
  return acc;
}
__device__ __forceinline__ float f2_lo(unsigned long long v)
{
  float a, b;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return a;
}

template <int V>
__global__ void __launch_bounds__(256) peak_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int iters,
                                                   unsigned long long* __restrict__ clk)
{
  uint32_t acc[kUnroll];
  uint32_t b[kUnroll];
#pragma unroll
  for (int i = 0; i < kUnroll; i++)
  {
    b[i]   = in[(threadIdx.x * 3 + i * 11 + 5) & 1023];
    acc[i] = in[(threadIdx.x + i * 37) & 1023];
  }
  float facc[kUnroll];
#pragma unroll
  for (int i = 0; i < kUnroll; i++) facc[i] = __uint_as_float(acc[i] >> 3);
  // packed state: slot pair j = slots 2j, 2j+1
  unsigned long long pacc[kUnroll / 2], pb[kUnroll / 2];
#pragma unroll
  for (int j = 0; j < kUnroll / 2; j++)
  {
    pacc[j] = f2_pack(facc[2 * j], facc[2 * j + 1]);
    pb[j]   = f2_pack(__uint_as_float(b[2 * j]), __uint_as_float(b[2 * j + 1]));
  }
  // number of leading slots on the ALU pipe in the packed mixes (the rest are FADD2 pairs)
  // (the older mixed variants V_VABSDIFF_FADD / V_VABSDIFF_FADD2 take their FP minuend from an ALU slot, which makes the
  // difference loop-invariant: ptxas hoists it and they run ONE FADD per pixel - kept for continuity, not used for peaks)
  constexpr int kAluSlots = (V == V_VABSDIFF_FADD2P_88 || V == V_MIX_FADD_88) ? 8 : V == V_VABSDIFF_FADD2P_610 ? 6
                          : (V == V_VABSDIFF_FADD2P_106 || V == V_MIX_FADD_106) ? 10
                          : (V == V_VABSDIFF_FADD2P_124 || V == V_MIX_FADD_124) ? 12 : 0;
  constexpr bool kPlainMix = V == V_MIX_FADD_124 || V == V_MIX_FADD_106 || V == V_MIX_FADD_88;
  constexpr bool kPacked = V >= V_FADD2P && !kPlainMix;

  unsigned long long g0 = 0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < iters; it++)
  {
    // operand x is another accumulator (static index), so nothing is loop-invariant and every chain stays
    // independent of its own previous result only through acc[i] (16 chains per thread).
#pragma unroll
    for (int i = 0; i < kUnroll; i++)
    {
      if (kPlainMix)
      {   // honest scalar mix: ALU slots take their operand from ALU slots, FP slots (2 FADD per pixel) from FP slots
        if (i < kAluSlots) acc[i] = __usad(acc[(i + 3) % kAluSlots], b[i], acc[i]);
        else
        {
          const int ix = kAluSlots + ((i - kAluSlots + 3) % (kUnroll - kAluSlots));
          float d = __fsub_rn(facc[ix], __uint_as_float(b[i]));
          facc[i] = __fadd_rn(facc[i], fabsf(d));
        }
        continue;
      }
      if (kPacked)
      {
        if (i < kAluSlots) acc[i] = __usad(acc[(i + 3) % (kAluSlots > 0 ? kAluSlots : 1)], b[i], acc[i]);
        else if ((i & 1) == 0)
        {
          const int j = i / 2, jx = (kAluSlots / 2) + ((j - kAluSlots / 2 + 3) % ((kUnroll - kAluSlots) / 2));
          pacc[j] = f2_absdiff_acc(pacc[j], f2_lo(pacc[jx]), pb[j]);
        }
        continue;
      }
      const uint32_t x = acc[(i + 5) & (kUnroll - 1)];
      const float    fx = facc[(i + 5) & (kUnroll - 1)];
      if (V == V_VABSDIFF) acc[i] = __usad(x, b[i], acc[i]);
      if (V == V_IADD3) acc[i] = acc[i] + x + b[i];
      if (V == V_IMAD) acc[i] = x * b[i] + acc[i];
      if (V == V_LOP3) acc[i] = (acc[i] & x) ^ b[i];
      if (V == V_PRMT) acc[i] = __byte_perm(acc[i], x, b[i]);
      if (V == V_VABSDIFF_IMAD)
      {
        if (i & 1) acc[i] = __usad(x, b[i], acc[i]);
        else       acc[i] = x * b[i] + acc[i];
      }
      if (V == V_FADD_ABS)
      {
        // |a-b| accumulated in FP32 on raw integer bit patterns (denormals: exact fixed point, no .ftz)
        float d = __fsub_rn(fx, __uint_as_float(b[i]));
        facc[i] = __fadd_rn(facc[i], fabsf(d));
      }
      if (V == V_VABSDIFF_FADD)
      {   // 2 px on the ALU pipe per 1 px on the FP pipe (1 + 1 + 2 instructions per 4 slots)
        if ((i & 3) < 2) acc[i] = __usad(x, b[i], acc[i]);
        if ((i & 3) == 3)
        {
          float d = __fsub_rn(fx, __uint_as_float(b[i]));
          facc[i] = __fadd_rn(facc[i], fabsf(d));
        }
      }
      if (V == V_VABSDIFF_FADD2)
      {   // 1 px ALU : 1 px FP  (1 + 2 instructions per 2 slots)
        if (i & 1) acc[i] = __usad(x, b[i], acc[i]);
        else
        {
          float d = __fsub_rn(fx, __uint_as_float(b[i]));
          facc[i] = __fadd_rn(facc[i], fabsf(d));
        }
      }
      if (V == V_VIADD16X2) acc[i] = __vadd2(x, b[i]) ^ acc[i];
      if (V == V_VIADDMNMX16X2) acc[i] = __viaddmax_s16x2(x, b[i], acc[i]);
      if (V == V_VABSDIFF4) acc[i] = __vsadu4(x, b[i]) + acc[i];
      if (V == V_DP2A) acc[i] = (uint32_t) __dp2a_lo((int) x, (int) b[i], (int) acc[i]);
      if (V == V_DP4A) acc[i] = (uint32_t) __dp4a((int) x, (int) b[i], (int) acc[i]);
    }
  }
  const long long t1 = clock64();
  uint32_t        r  = 0;
#pragma unroll
  for (int i = 0; i < kUnroll; i++) r += acc[i] + __float_as_uint(facc[i]);
#pragma unroll
  for (int j = 0; j < kUnroll / 2; j++) r += (uint32_t) pacc[j] + (uint32_t) (pacc[j] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0)
  {
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    clk[2 * blockIdx.x]     = (unsigned long long) (t1 - t0);   // SM cycles of this CTA's span
    clk[2 * blockIdx.x + 1] = g1 - g0;                          // nanoseconds of the same span
  }
}

// instructions per accumulator-slot per iteration (for the rate computation)
__host__ double instr_per_slot(int v)
{
  switch (v)
  {
    case V_FADD_ABS: return 2.0;
    case V_VABSDIFF_FADD: return (2 * 1.0 + 1 * 2.0) / 4.0;   // slots 0,1 usad; slot 2 idle; slot 3 = 2 FADD
    case V_VIADD16X2: return 2.0;   // VIADD.16x2 + LOP3
    case V_VABSDIFF_FADD2: return 1.5;
    case V_MIX_FADD_124: return (12 + 2 * 4) / 16.0;
    case V_MIX_FADD_106: return (10 + 2 * 6) / 16.0;
    case V_MIX_FADD_88: return (8 + 2 * 8) / 16.0;
    case V_VABSDIFF4: return 1.0;   // ptxas folds the add into VABSDIFF4.U8.ACC
    default: return 1.0;
  }
}

template <int V>
int run_variant(int iters, int sms, const uint32_t* din, uint32_t* dout, unsigned long long* dclk, double* laneRate,
                double* ms, double* mhzOut, cudaStream_t st)
{
  const int   grid = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  peak_kernel<V><<<grid, 256, 0, st>>>(din, dout, iters / 8 + 1, dclk);   // warm-up
  cudaEventRecord(e0, st);
  peak_kernel<V><<<grid, 256, 0, st>>>(din, dout, iters, dclk);
  cudaEventRecord(e1, st);
  cudaError_t err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) return -1;
  float t = 0;
  cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // SM clock = cycles / nanoseconds of CTA spans (averaged over the first 64 CTAs); the kernel's cycle count is
  // its event-timed duration at that clock.  (CTA spans themselves are useless for the rate: the scheduler runs
  // co-resident CTAs greedily, one after the other.)
  unsigned long long hclk[128];
  cudaMemcpy(hclk, dclk, sizeof(hclk), cudaMemcpyDeviceToHost);
  double cyc = 0, ns = 0;
  for (int i = 0; i < 64; i++)
  {
    cyc += (double) hclk[2 * i];
    ns += (double) hclk[2 * i + 1];
  }
  const double mhz           = ns > 0 ? cyc / ns * 1e3 : 0.0;
  const double kernelCycles  = (double) t * 1e-3 * mhz * 1e6;
  const double laneInstrPerSm = 8.0 * 256.0 * (double) iters * kUnroll * instr_per_slot(V);
  *laneRate = kernelCycles > 0 ? laneInstrPerSm / kernelCycles : 0.0;
  *ms       = t;
  *mhzOut   = mhz;
  return 0;
}

}   // namespace

extern "C" int vtmme_int_peak(int variant, int iters, double* laneInstrPerClkPerSm, double* ms, double* smClockMHz)
{
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return VTMME_ERR_CUDA;
  const int sms = prop.multiProcessorCount;
  uint32_t *din = nullptr, *dout = nullptr;
  unsigned long long* dclk = nullptr;
  cudaMalloc(&din, 1024 * 4);
  cudaMalloc(&dout, (size_t) sms * 8 * 256 * 4);
  cudaMalloc(&dclk, (size_t) sms * 8 * 16);
  uint32_t h[1024];
  for (int i = 0; i < 1024; i++) h[i] = (uint32_t) ((i * 2654435761u) >> 22);   // 10-bit values
  cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaStream_t st;
  cudaStreamCreate(&st);
  int rc = -1;
  switch (variant)
  {
#define CASE(V) case V: rc = run_variant<V>(iters, sms, din, dout, dclk, laneInstrPerClkPerSm, ms, smClockMHz, st); break;
    CASE(V_VABSDIFF) CASE(V_IADD3) CASE(V_IMAD) CASE(V_LOP3) CASE(V_PRMT) CASE(V_VABSDIFF_IMAD) CASE(V_FADD_ABS)
    CASE(V_VABSDIFF_FADD) CASE(V_VIADD16X2) CASE(V_VIADDMNMX16X2) CASE(V_VABSDIFF4) CASE(V_VABSDIFF_FADD2) CASE(V_DP2A) CASE(V_DP4A)
    CASE(V_FADD2P) CASE(V_VABSDIFF_FADD2P_88) CASE(V_VABSDIFF_FADD2P_610) CASE(V_VABSDIFF_FADD2P_106)
    CASE(V_VABSDIFF_FADD2P_124) CASE(V_MIX_FADD_124) CASE(V_MIX_FADD_106) CASE(V_MIX_FADD_88)
#undef CASE
    default: rc = -2;
  }
  cudaStreamDestroy(st);
  cudaFree(din);
  cudaFree(dout);
  cudaFree(dclk);
  return rc == 0 ? VTMME_OK : VTMME_ERR_CUDA;
}
