// variants of the all-to-all hand-over (see handover_bench.cu) + a two-CTA ping-pong for the uncontended latency
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint4 ld_poll4(const uint32_t *p) {
  uint4 r; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory"); return r;
}
__device__ __forceinline__ uint32_t ld_poll1(const uint32_t *p) { uint32_t r; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory"); return r; }
__device__ __forceinline__ void st_unit(uint32_t *p, uint32_t u) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(u) : "memory"); }
__device__ __forceinline__ bool ok4(const uint4 &u, uint32_t t) { return ((u.x & 0xFFFF) == t) & ((u.y & 0xFFFF) == t) & ((u.z & 0xFFFF) == t) & ((u.w & 0xFFFF) == t); }
// mode 0: thread t polls units [8t, 8t+8) as two strided 16-byte loads (the kernel's current layout)
// mode 1: thread t polls [4t,4t+4) and [K/2+4t, +4): both loads fully coalesced
// mode 2: mode 1, but only ONE warp group polls (threads < K/8) while a second copy of the pollers (threads K/8 .. K/4) polls
//         half a period later (two polls in flight per datum)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint32_t *buf0, uint32_t *buf1, int K, int ustride, int R, int phases, long long *out) {
  const int tid = threadIdx.x, bid = blockIdx.x, grid = gridDim.x;
  const int r0 = (int)((long long)K * bid / grid), r1 = (int)((long long)K * (bid + 1) / grid);
  __shared__ float sm[4096]; __shared__ volatile int s_got;
  unsigned long long g0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
  for (int ph = 1; ph <= phases; ++ph) {
    uint32_t *in = (ph & 1) ? buf0 : buf1, *outb = (ph & 1) ? buf1 : buf0;
    const uint32_t tag_in = (uint32_t)(ph - 1) & 0xFFFF, tag = (uint32_t)ph & 0xFFFF;
    const int T = K / 8;
    if (ph > 1 && tid < T) {
      const uint32_t *base = in + (size_t)(bid % R) * ustride;
      const uint32_t *pa = MODE == 0 ? base + tid * 8 : base + tid * 4, *pb = MODE == 0 ? pa + 4 : base + K / 2 + tid * 4;
      uint4 a, b;
      if (MODE == 2) {         // two polls in flight, half a round trip apart
        uint4 a2, b2;
        a = ld_poll4(pa); b = ld_poll4(pb);
        for (;;) {
          __nanosleep(100);
          a2 = ld_poll4(pa); b2 = ld_poll4(pb);
          if (ok4(a, tag_in) & ok4(b, tag_in)) break;
          __nanosleep(100);
          a = ld_poll4(pa); b = ld_poll4(pb);
          if (ok4(a2, tag_in) & ok4(b2, tag_in)) { a = a2; b = b2; break; }
        }
      } else
      for (;;) { a = ld_poll4(pa); b = ld_poll4(pb); if (ok4(a, tag_in) & ok4(b, tag_in)) break; }
      sm[tid * 8] = __uint_as_float(a.x & 0xFFFF0000u) + __uint_as_float(b.w & 0xFFFF0000u);
    }
    __syncthreads();
    for (int r = r0 + tid; r < r1; r += blockDim.x) {
      const uint32_t u = (__float_as_uint(sm[(r * 8) % K]) & 0xFFFF0000u) | tag;
      for (int q = 0; q < R; ++q) st_unit(outb + (size_t)q * ustride + r, u);
    }
    __syncthreads();
  }
  unsigned long long g1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
  if (tid == 0 && bid == 0) out[0] = (long long)(g1 - g0);
}
__global__ void pingpong(uint32_t *f, int iters, long long *out) {
  if (threadIdx.x) return;
  unsigned long long g0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
  const int me = blockIdx.x;
  for (int i = 1; i <= iters; ++i) {
    if (me == 0) { st_unit(f, i); while (ld_poll1(f + 64) != (uint32_t)i) {} }
    else { while (ld_poll1(f) != (uint32_t)i) {} st_unit(f + 64, i); }
  }
  unsigned long long g1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
  if (me == 0) out[0] = (long long)(g1 - g0);
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int ustride = 4096 + 64, phases = 2000;
  uint32_t *b0, *b1; long long *out;
  cudaMalloc(&b0, (size_t)16 * ustride * 4); cudaMalloc(&b1, (size_t)16 * ustride * 4); cudaMalloc(&out, 64);
  long long h;
  { int it = 2000; void *args[] = {&b0, &it, &out}; cudaMemset(b0, 0, 1024);
    for (int grid : {2, 148}) { cudaMemset(b0, 0, 1024); cudaLaunchCooperativeKernel((void *)pingpong, dim3(grid > 2 ? 2 : 2), dim3(32), args, 0, 0); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); printf("ping-pong round trip (2 CTAs, 1 unit each way): %.3f us\n", h / 1e3 / it); } }
  for (int mode = 0; mode < 3; ++mode) for (int K : {1024, 3072}) for (int R : {1}) for (int grid : {148}) {
    cudaMemset(b0, 0xFF, (size_t)16 * ustride * 4); cudaMemset(b1, 0xFF, (size_t)16 * ustride * 4);
    int Kk = K, us = ustride, ph = phases, rr = R;
    void *args[] = {&b0, &b1, &Kk, &us, &rr, &ph, &out};
    cudaError_t e = cudaLaunchCooperativeKernel(mode == 0 ? (void *)k<0> : mode == 1 ? (void *)k<1> : (void *)k<2>, dim3(grid), dim3(512), args, 0, 0);
    cudaError_t e2 = cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("mode %d K=%d R=%2d grid=%3d: %.3f us per hand-over  %s %s\n", mode, K, R, grid, h / 1e3 / phases, e ? cudaGetErrorString(e) : "", e2 ? cudaGetErrorString(e2) : "");
  }
  return 0;
}
