// floor of one all-to-all hand-over through tagged units: 148 co-resident CTAs, each polls the whole K-unit vector
// (replica bid % R), then writes its own slice of the next vector (R replicas).  No compute in between.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define R 4
__device__ __forceinline__ uint4 ld_poll4(const uint32_t *p) {
  uint4 r; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory"); return r;
}
__device__ __forceinline__ void st_unit(uint32_t *p, uint32_t u) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(u) : "memory"); }
__device__ __forceinline__ bool ok4(const uint4 &u, uint32_t t) { return ((u.x & 0xFFFF) == t) & ((u.y & 0xFFFF) == t) & ((u.z & 0xFFFF) == t) & ((u.w & 0xFFFF) == t); }
__global__ void __launch_bounds__(512, 1) k(uint32_t *buf0, uint32_t *buf1, int K, int ustride, int phases, int sleep_ns, long long *out) {
  const int tid = threadIdx.x, bid = blockIdx.x, grid = gridDim.x;
  const int r0 = (int)((long long)K * bid / grid), r1 = (int)((long long)K * (bid + 1) / grid);
  __shared__ float sm[4096];
  long long t0 = clock64();
  unsigned long long g0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
  for (int ph = 1; ph <= phases; ++ph) {
    uint32_t *in = (ph & 1) ? buf0 : buf1, *outb = (ph & 1) ? buf1 : buf0;
    const uint32_t tag_in = (uint32_t)(ph - 1) & 0xFFFF, tag = (uint32_t)ph & 0xFFFF;
    if (ph > 1 && tid * 8 < K) {
      const uint32_t *p = in + (size_t)(bid % R) * ustride + tid * 8;
      uint4 a, b;
      for (;;) { a = ld_poll4(p); b = ld_poll4(p + 4); if (ok4(a, tag_in) & ok4(b, tag_in)) break; if (sleep_ns) __nanosleep(sleep_ns); }
      sm[tid * 8] = __uint_as_float(a.x & 0xFFFF0000u) + __uint_as_float(b.w & 0xFFFF0000u);
    }
    __syncthreads();
    for (int r = r0 + tid; r < r1; r += blockDim.x) {
      const uint32_t u = (__float_as_uint(sm[(r * 8) % K]) & 0xFFFF0000u) | tag;
#pragma unroll
      for (int q = 0; q < R; ++q) st_unit(outb + (size_t)q * ustride + r, u);
    }
    __syncthreads();
  }
  unsigned long long g1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
  if (tid == 0 && bid == 0) { out[0] = clock64() - t0; out[1] = (long long)(g1 - g0); }
}
int main(int argc, char **argv) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int ustride = 4096 + 64, phases = 2000;
  uint32_t *b0, *b1; long long *out;
  cudaMalloc(&b0, (size_t)R * ustride * 4); cudaMalloc(&b1, (size_t)R * ustride * 4); cudaMalloc(&out, 64);
  for (int K : {1024, 3072}) for (int sleep_ns : {0, 20, 100}) {
    cudaMemset(b0, 0xFF, (size_t)R * ustride * 4); cudaMemset(b1, 0xFF, (size_t)R * ustride * 4);
    int Kk = K, us = ustride, ph = phases, sl = sleep_ns;
    void *args[] = {&b0, &b1, &Kk, &us, &ph, &sl, &out};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)k, dim3(sms), dim3(512), args, 0, 0);
    long long h[2]; cudaError_t e2 = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("K=%d sleep=%d ns: %.3f us per hand-over (%lld cycles)  %s %s\n", K, sleep_ns, h[1] / 1e3 / phases, h[0] / phases, cudaGetErrorString(e), cudaGetErrorString(e2));
  }
  return 0;
}
