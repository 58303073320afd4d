// standalone device check of the bitonic network used by the sort-based sampler (sampler.cuh)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../fish_tts_b200/csrc/sampler.cuh"
using namespace da;
template <int E, int NT, class B>
__global__ void k_sort(const uint32_t *in, uint32_t *out, int n) {
  extern __shared__ uint32_t sm[];
  if (threadIdx.x >= NT) return;
  uint32_t a[E];
#pragma unroll
  for (int i = 0; i < E; ++i) { int e = threadIdx.x * E + i; a[i] = e < n ? in[e] : 0xFFFFFFFFu; }
  bitonic_sort_u32<E, NT, B>(a, sm);
#pragma unroll
  for (int i = 0; i < E; ++i) out[threadIdx.x * E + i] = a[i];
}
template <int E, int NT, class B> int run(int n, int nthreads) {
  std::vector<uint32_t> h(n); for (auto &v : h) v = (uint32_t)rand() * 2654435761u;
  uint32_t *din, *dout; cudaMalloc(&din, n * 4); cudaMalloc(&dout, E * NT * 4);
  cudaMemcpy(din, h.data(), n * 4, cudaMemcpyHostToDevice);
  k_sort<E, NT, B><<<1, nthreads, E * NT * 4>>>(din, dout, n);
  std::vector<uint32_t> o(E * NT); cudaError_t e = cudaMemcpy(o.data(), dout, E * NT * 4, cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  int bad = 0; for (int i = 0; i < n; ++i) bad += o[i] != h[i];
  printf("E=%d NT=%d n=%d: %s (%d wrong) %s\n", E, NT, n, bad ? "FAIL" : "ok", bad, cudaGetErrorString(e));
  return bad;
}
int main() {
  int bad = 0;
  bad += run<8, 128, BlockNamed<2, 128>>(256, 544);
  bad += run<8, 128, BlockNamed<2, 128>>(1024, 544);
  bad += run<8, 512, BlockNamed<1, 512>>(256, 544);
  bad += run<8, 512, BlockNamed<1, 512>>(640, 544);
  bad += run<8, 512, BlockNamed<1, 512>>(4096, 544);
  return bad != 0;
}
