#include <cstdio>
#include <cstdint>
__global__ void k(const uint32_t* in, uint32_t* hi, uint32_t* lo, uint32_t* hi2, uint32_t* lo2, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float wf = __uint_as_float(in[i]);
  float h = __fmul_rn(wf, 0.000244140625f);
  hi[i] = __float_as_uint(h);
  lo[i] = __float_as_uint(__fmaf_rn(h, -4096.0f, wf));
  // alternative: add-based split with a magic constant: (w + M) - M rounds w to a multiple of 4096 when M = 2^23 * 4096 ... in denormal domain M = bits 4096*2^23?
  float m = __uint_as_float(0x0b800000u);  // 2^-104 : ulp = 2^-127?  (just an experiment)
  float t = __fadd_rn(wf, m);
  hi2[i] = __float_as_uint(t);
  lo2[i] = __float_as_uint(__fsub_rn(t, m));
}
int main()
{
  const int n = 8;
  uint32_t h[n] = { 5u | (7u << 12), 1023u | (1023u << 12), 0u | (1u << 12), 1u, 2047u | (3u << 12), 512u | (512u << 12), 100u | (900u << 12), 4095u };
  uint32_t *d, *a, *b, *c, *e;
  cudaMalloc(&d, n * 4); cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&c, n * 4); cudaMalloc(&e, n * 4);
  cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
  k<<<1, 32>>>(d, a, b, c, e, n);
  uint32_t ra[n], rb[n], rc[n], re[n];
  cudaMemcpy(ra, a, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(rb, b, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(rc, c, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(re, e, n * 4, cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; i++) printf("w=%08x want hi=%u lo=%u | fmul hi=%08x fma lo=%08x | t=%08x t-m=%08x\n", h[i], h[i] >> 12, h[i] & 4095, ra[i], rb[i], rc[i], re[i]);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
