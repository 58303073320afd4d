#include <cstdio>
#include "../../fish_tts_b200/csrc/sampler.cuh"
using namespace da;
typedef BlockNamed<2, 128> G4;
__global__ void __launch_bounds__(544, 1) k(long long *cyc, uint32_t *out, int iters) {
  extern __shared__ uint32_t sm[];
  if (threadIdx.x >= 128) return;
  uint32_t a[8];
  long long t = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (threadIdx.x * 8 + i + it) * 2654435761u;
    long long c0 = clock64();
    bitonic_sort_u32<8, 128, G4>(a, sm);
    t += clock64() - c0;
    out[threadIdx.x] = a[0] ^ a[7];
  }
  // pweight cost
  long long c0 = clock64(); unsigned long long s = 0;
  for (int it = 0; it < iters; ++it) s += pweight(-0.5f - it * 0.001f, 0.25f, 300.f);
  long long tp = clock64() - c0;
  // noise cost
  c0 = clock64(); float f = 0;
  for (int it = 0; it < iters; ++it) f += exp1_noise(1234, 5, 3, it + threadIdx.x);
  long long tn = clock64() - c0;
  if (threadIdx.x == 0) { cyc[0] = t / iters; cyc[1] = tp / iters; cyc[2] = tn / iters; out[200] = (uint32_t)s + (uint32_t)f; }
}
int main() {
  long long *c; uint32_t *o; cudaMalloc(&c, 64); cudaMalloc(&o, 4096);
  k<<<1, 544, 8192>>>(c, o, 200);
  long long h[3]; cudaError_t e = cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost);
  printf("bitonic<8,128>: %lld cyc | pweight: %lld cyc | exp1_noise: %lld cyc  %s\n", h[0], h[1], h[2], cudaGetErrorString(e));
}
