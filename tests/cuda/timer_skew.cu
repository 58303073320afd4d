// is %globaltimer consistent across SMs?  All CTAs of a lock-step all-to-all loop stamp the same phase; the spread of the
// stamps bounds skew + hand-over latency.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <vector>
__device__ __forceinline__ uint32_t ld_poll1(const uint32_t *p) { uint32_t r; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory"); return r; }
__device__ __forceinline__ void st_unit(uint32_t *p, uint32_t u) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(u) : "memory"); }
__global__ void k(uint32_t *flags, unsigned long long *stamps, int phases) {
  const int bid = blockIdx.x, grid = gridDim.x, tid = threadIdx.x;
  for (int ph = 1; ph <= phases; ++ph) {
    if (tid == 0) st_unit(flags + bid * 32, ph);                 // my arrival
    if (tid < grid) while (ld_poll1(flags + tid * 32) < (uint32_t)ph) {}
    __syncthreads();
    if (tid == 0) { unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); stamps[(size_t)ph * grid + bid] = g; }
  }
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int phases = 200;
  uint32_t *f; unsigned long long *s; cudaMalloc(&f, sms * 128); cudaMalloc(&s, (size_t)(phases + 1) * sms * 8); cudaMemset(f, 0, sms * 128);
  int ph = phases; void *args[] = {&f, &s, &ph};
  cudaLaunchCooperativeKernel((void *)k, dim3(sms), dim3(256), args, 0, 0);
  std::vector<unsigned long long> h((size_t)(phases + 1) * sms); cudaError_t e = cudaMemcpy(h.data(), s, h.size() * 8, cudaMemcpyDeviceToHost);
  double worst = 0, mean = 0; 
  for (int p = 50; p <= phases; ++p) { auto b = h.begin() + (size_t)p * sms; auto mm = std::minmax_element(b, b + sms); double sp = (double)(*mm.second - *mm.first); worst = std::max(worst, sp); mean += sp; }
  printf("grid barrier via flags: spread of %%globaltimer stamps across %d CTAs: mean %.0f ns, worst %.0f ns; period %.0f ns  %s\n", sms, mean / (phases - 49), worst,
         (double)(h[(size_t)phases * sms] - h[(size_t)50 * sms]) / (phases - 50), cudaGetErrorString(e));
  return 0;
}
