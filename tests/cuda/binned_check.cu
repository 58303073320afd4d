// standalone exactness check: sample_binned() against sample_sorted() on random logit vectors (token and nucleus size must agree)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define DA_SAMPLER_TIMING 1
#include "../../fish_tts_b200/csrc/sampler.cuh"
using namespace da;
template <int E, int NT, class G>
__global__ void __launch_bounds__(544, 1) k_check(const uint16_t *logits, int V, float top_p, float s_factor, int all_present, int which, uint32_t *tok_out, int *nucleus, long long *cyc) {
  extern __shared__ __align__(16) unsigned char sm[];
  unsigned long long *scr = reinterpret_cast<unsigned long long *>(sm);
  const int tid = threadIdx.x, w = tid >> 5;
  if (w >= NT / 32) return;
  NoiseSrc ns = {nullptr, 1234ull, 7u};
  uint32_t it8[E]; Red r = {0ull, 0, -1};
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int e = tid * E + i; it8[i] = 0xFFFFFFFFu;
    if (e < V) { const uint32_t key = bf16_key(logits[e]); it8[i] = ((0xFFFFu - key) << 16) | (uint32_t)e; r.m = max(r.m, (int)key); }
  }
  SampleParams spm; int par = 0;
  r = block_reduce<G>(r, scr, par);
  spm.m = bits2f(key_bf16((uint32_t)r.m));
  Red es = {0ull, 0, -1};
#pragma unroll
  for (int i = 0; i < E; ++i) if (it8[i] != 0xFFFFFFFFu) es.s += (unsigned long long)(expf(bits2f(key_bf16(0xFFFFu - (it8[i] >> 16))) - spm.m) * DA_FIX2_SCALE);
  es = block_reduce<G>(es, scr, par);
  spm.S = __ull2float_rn(es.s) * (1.0f / DA_FIX2_SCALE) * s_factor;
  spm.T_bf = 0.69921875f; spm.c_max = cmax_from_top_p(top_p);
  G::sync();
  const long long c0 = clock64();
  uint32_t tok;
  if (which == 0) tok = sample_sorted<E, NT, G>(it8, (uint32_t)V, all_present != 0, nullptr, spm, ns, 3u, 0ll, nucleus, reinterpret_cast<uint32_t *>(scr + 256), scr);
  else tok = sample_binned<E, NT, G>(it8, (uint32_t)V, all_present != 0, nullptr, spm, ns, 3u, 0ll, nucleus, reinterpret_cast<uint32_t *>(scr + 256), scr);
  const long long c1 = clock64();
  if (tid == 0) { tok_out[0] = tok; cyc[0] = c1 - c0; }
}
static uint16_t f2b(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFF + ((u >> 16) & 1); return (uint16_t)(u >> 16); }
int main() {
  uint16_t *d; uint32_t *t; int *n; long long *c;
  cudaMalloc(&d, 4096 * 2); cudaMalloc(&t, 16); cudaMalloc(&n, 16); cudaMalloc(&c, 16);
  const float tops[] = {1e-9f, 0.05f, 0.3f, 0.8f, 0.95f, 1.0f};
  int bad = 0, total = 0; long long cyc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  for (int big = 0; big < 3; ++big) {      // 0: 1024 items, 128 threads x 8; 1: 4096 items, 512 x 8; 2: 1024 items, 512 x 2
    const int V = big == 1 ? 4096 : 1024;
    for (int dist = 0; dist < 8; ++dist) for (int seed = 0; seed < 6; ++seed) for (float top_p : tops) for (int ap = 0; ap < 2; ++ap) {
      srand(1000 * dist + seed);
      std::vector<uint16_t> h(V);
      for (int i = 0; i < V; ++i) {
        const float u = (float)rand() / RAND_MAX;
        float f;
        switch (dist) {
          case 0: f = (u - 0.5f) * 6.0f; break;                       // flat, both signs
          case 1: f = u * 3.0f; break;                                 // positive, narrow
          case 2: f = (i % 97 == 0) ? 9.0f + u : u * 2.0f; break;      // a few peaks
          case 3: f = floorf(u * 8.0f) * 0.25f; break;                 // heavy ties
          case 4: f = 1.5f; break;                                     // all equal
          case 5: f = (i == 123) ? 20.0f : (u - 0.5f) * 4.0f; break;   // one dominant
          case 6: f = -30.0f * u; break;                               // long negative tail
          default: f = (u < 0.5f) ? 2.0f + 0.001f * u : -50.0f; break; // half in one bin, half far away
        }
        h[i] = f2b(f);
      }
      cudaMemcpy(d, h.data(), V * 2, cudaMemcpyHostToDevice);
      uint32_t tk[2]; int nk[2]; long long cy[2];
      for (int which = 0; which < 2; ++which) {
        cudaMemset(n, 0xFF, 4);
        const float sf = ap ? 1.0f : 3.0f;      // candidates of a larger vocabulary: S covers more than the items
        if (big == 1) k_check<8, 512, BlockNamed<2, 512>><<<1, 544, 32768>>>(d, V, top_p, sf, ap, which, t, n, c);
        else if (big == 2) k_check<2, 512, BlockNamed<2, 512>><<<1, 544, 32768>>>(d, V, top_p, sf, ap, which, t, n, c);
        else k_check<8, 128, BlockNamed<2, 128>><<<1, 544, 32768>>>(d, V, top_p, sf, ap, which, t, n, c);
        cudaError_t e = cudaMemcpy(&tk[which], t, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&nk[which], n, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cy[which], c, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cyc[big][which] += cy[which];
      }
      if (seed == 0 && ap == 1 && (top_p == 0.8f)) {
        long long ts[16]; cudaMemcpyFromSymbol(ts, g_sampler_t, sizeof(ts));
        printf("V %d dist %d binned sections (cycles): hist %lld | scan+cut %lld | list %lld | rank %lld | keep+S2 %lld | race %lld | argbest %lld | total %lld nucleus %d\n", V, dist, ts[1]-ts[0], ts[2]-ts[1], ts[3]-ts[2], ts[4]-ts[3], ts[5]-ts[4], ts[6]-ts[5], ts[7]-ts[6], cy[1], nk[1]);
      }
      ++total;
      if (tk[0] != tk[1] || nk[0] != nk[1]) { if (++bad <= 20) printf("MISMATCH V %d dist %d seed %d top_p %g all_present %d: sorted tok %u nucleus %d | binned tok %u nucleus %d\n", V, dist, seed, top_p, ap, tk[0], nk[0], tk[1], nk[1]); }
    }
  }
  printf("cases %d mismatches %d | mean cycles 1024 items / 128 threads: sorted %lld binned %lld | 4096 / 512: sorted %lld binned %lld | 1024 / 512: sorted %lld binned %lld\n", total, bad,
         cyc[0][0] / (total / 3), cyc[0][1] / (total / 3), cyc[1][0] / (total / 3), cyc[1][1] / (total / 3), cyc[2][0] / (total / 3), cyc[2][1] / (total / 3));
  return bad != 0;
}
