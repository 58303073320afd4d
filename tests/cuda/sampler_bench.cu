// standalone timing of the fast-head sampling block of mega.cuh (one CTA, warps 0..3), in isolation
#include <cstdio>
#include <cstdlib>
#include <vector>
#define DA_SAMPLER_TIMING 1
#include "../../fish_tts_b200/csrc/sampler.cuh"
using namespace da;
template <int E, int NT, class G4>
__global__ void __launch_bounds__(544, 1) k_bench(const uint16_t *logits, int V, int iters, long long *cyc, uint32_t *tok_out, int *nucleus) {
  extern __shared__ __align__(16) unsigned char sm[];
  unsigned long long *scr = reinterpret_cast<unsigned long long *>(sm);
  const int tid = threadIdx.x, w = tid >> 5;
  if (w >= NT / 32) return;
  NoiseSrc ns = {nullptr, 1234ull, 7u};
  long long t_sort = 0, t_all = 0, t_ms = 0;
  uint32_t tok = 0;
  for (int it = 0; it < iters; ++it) {
    long long c0 = clock64();
    uint32_t it8[E]; Red r = {0ull, 0, -1};
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int e = tid * E + i; it8[i] = 0xFFFFFFFFu;
      if (e < V) { const uint32_t key = bf16_key(logits[(e + it) % V]); it8[i] = ((0xFFFFu - key) << 16) | (uint32_t)e; r.m = max(r.m, (int)key); }
    }
    SampleParams spm; int par = 0;
    r = block_reduce<G4>(r, scr, par);
    spm.m = bits2f(key_bf16((uint32_t)r.m));
    Red es = {0ull, 0, -1};
#pragma unroll
    for (int i = 0; i < E; ++i) if (it8[i] != 0xFFFFFFFFu) es.s += (unsigned long long)(expf(bits2f(key_bf16(0xFFFFu - (it8[i] >> 16))) - spm.m) * DA_FIX2_SCALE);
    es = block_reduce<G4>(es, scr, par);
    spm.S = __ull2float_rn(es.s) * (1.0f / DA_FIX2_SCALE);
    spm.T_bf = 0.69921875f; spm.c_max = cmax_from_top_p(0.8f);
    G4::sync();
    long long c1 = clock64();
    tok = sample_sorted<E, NT, G4>(it8, (uint32_t)V, true, nullptr, spm, ns, 3u, 0ll, nucleus, reinterpret_cast<uint32_t *>(scr + 256), scr);
    long long c2 = clock64();
    t_ms += c1 - c0; t_sort += c2 - c1; t_all += c2 - c0;
  }
  if (tid == 0) { cyc[0] = t_ms / iters; cyc[1] = t_sort / iters; cyc[2] = t_all / iters; tok_out[0] = tok; }
}
int main() {
  const int V = 1024, iters = 200;
  std::vector<uint16_t> h(V);
  for (auto &v : h) { float f = ((float)rand() / RAND_MAX - 0.5f) * 6.0f; uint32_t u; memcpy(&u, &f, 4); v = (uint16_t)(u >> 16); }
  uint16_t *d; long long *c; uint32_t *t; int *n;
  cudaMalloc(&d, V * 2); cudaMalloc(&c, 64); cudaMalloc(&t, 16); cudaMalloc(&n, 16);
  cudaMemcpy(d, h.data(), V * 2, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 3; ++variant) {
  if (variant == 0) k_bench<8, 128, BlockNamed<2, 128>><<<1, 544, 16384>>>(d, V, iters, c, t, n);
  else if (variant == 1) k_bench<4, 256, BlockNamed<2, 256>><<<1, 544, 16384>>>(d, V, iters, c, t, n);
  else k_bench<2, 512, BlockNamed<2, 512>><<<1, 544, 16384>>>(d, V, iters, c, t, n);
  long long hc[3]; uint32_t ht; int hn;
  cudaError_t e = cudaMemcpy(hc, c, 24, cudaMemcpyDeviceToHost); cudaMemcpy(&ht, t, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hn, n, 4, cudaMemcpyDeviceToHost);
  long long ts[16]; cudaMemcpyFromSymbol(ts, g_sampler_t, sizeof(ts));
  printf("sections: sort %lld | pweight+cum %lld | scan+count %lld | e2 %lld | S2 reduce %lld | noise/race %lld | argbest %lld\n", ts[1]-ts[0], ts[2]-ts[1], ts[3]-ts[2], ts[4]-ts[3], ts[5]-ts[4], ts[6]-ts[5], ts[7]-ts[6]);
  printf("variant %d fast-head sampler alone: m,S %lld cyc | sample_sorted %lld cyc | total %lld cyc  (tok %u nucleus %d) %s\n", variant, hc[0], hc[1], hc[2], ht, hn, cudaGetErrorString(e));
  }
  return 0;
}
