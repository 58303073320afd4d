// the compute section of a GEMV phase of mega.cuh in isolation: tiles already in shared memory, 16 warps, units -> partial
// sums -> "last finisher folds".  Reports cycles from the barrier after staging to the moment every warp has left the loop.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../fish_tts_b200/csrc/mega.cuh"
using namespace da;
__global__ void __launch_bounds__(DA_M_THREADS, 1) k(int rows, int K, int epi, int reps, uint32_t *out, long long *cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (w >= DA_M_CWARPS) return;
  const int nchunk = K >> 7, nt = (rows + 15) >> 4;
  const uint32_t RS = row_stride(K);
  bf16 *xb = reinterpret_cast<bf16 *>(sm);
  float *part = reinterpret_cast<float *>(sm + 2 * K);
  volatile int *pcnt = reinterpret_cast<volatile int *>(part + DA_M_PT * nchunk * 16), *pgen = pcnt + DA_M_PT;
  float *resv = reinterpret_cast<float *>(const_cast<int *>(pgen + DA_M_PT));
  unsigned char *tiles = sm + ((2 * K + DA_M_PT * nchunk * 64 + 64 + 4096 * 4 + 127) & ~127);
  for (int i = tid; i < K; i += DA_M_CTHREADS) xb[i] = f2bf(0.01f * (float)((i * 37) % 101 - 50));
  for (int i = tid; i < 4096; i += DA_M_CTHREADS) resv[i] = 0.5f;
  for (int i = tid; i < (int)(nt * 16 * RS / 2); i += DA_M_CTHREADS) reinterpret_cast<bf16 *>(tiles)[i] = f2bf(0.001f * (float)((i * 13) % 97 - 48));
  if (tid < 2 * DA_M_PT) pcnt[tid] = 0;
  cbar();
  int gen_base[DA_M_PT] = {0, 0, 0, 0};
  long long total = 0, t_mma = 0, t_sig = 0, t_fold = 0;
  const uint32_t *xw = reinterpret_cast<const uint32_t *>(xb);
  for (int rep = 0; rep < reps; ++rep) {
    const uint32_t tag = (uint32_t)rep & 0xFFFF;
    cbar();
    const long long c0 = clock64();
    int u = w;
    for (int t = 0; t < nt; ++t) {
      const int n = min(16, rows - 16 * t);
      const uint32_t at = (uint32_t)t * 16 * RS;
      if (u >= (t + 1) * nchunk) continue;
      const int slot = t & (DA_M_PT - 1);
      int gb = gen_base[0];
#pragma unroll
      for (int i = 1; i < DA_M_PT; ++i) if (slot == i) gb = gen_base[i];
      const int gen_need = gb + t / DA_M_PT;
      float *pslot = part + (size_t)slot * nchunk * 16;
      for (; u < (t + 1) * nchunk; u += DA_M_CWARPS) {
        const int c = u - t * nchunk;
        float v_lo, v_hi;
        const long long m0 = clock64();
        mma_chunk(smem_u32(tiles + at), RS, n, xw, c, lane, v_lo, v_hi);
        const long long m1 = clock64();
        if (t >= DA_M_PT) { while (pgen[slot] != gen_need) { __nanosleep(20); } }
        if ((lane & 3) == 0) { pslot[c * 16 + (lane >> 2)] = v_lo; pslot[c * 16 + 8 + (lane >> 2)] = v_hi; }
        __syncwarp();
        int last = 0;
        if (lane == 31) last = (atom_add_acq_rel_cta(&pcnt[slot], 1) == nchunk - 1);
        last = __shfl_sync(0xffffffffu, last, 31);
        const long long m2 = clock64();
        t_mma += m1 - m0; t_sig += m2 - m1;
        if (last) {
          const int r = lane & 15, row = 16 * t + r;
          const bool live = lane < 16 && r < n;
          float v = 0.f;
          for (int c8 = 0; c8 < nchunk; c8 += 8) {
            float pv[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) pv[cc] = (c8 + cc < nchunk) ? pslot[(c8 + cc) * 16 + r] : 0.f;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) if (c8 + cc < nchunk) v += pv[cc];
          }
          __syncwarp();
          if (lane == 31) { pcnt[slot] = 0; __threadfence_block(); pgen[slot] = gen_need + 1; }
          if (epi == ME_STORE) { if (live) st_unit(out + row, make_unit(v, tag)); }
          else if (epi == ME_RESIDUAL) { if (live) st_unit(out + row, make_unit(resv[row] + rbf(v), tag)); }
          else {
            const float up = __shfl_down_sync(0xffffffffu, v, 1);
            if (live && !(r & 1)) {
              const float gg = rbf(v), uu = rbf(up);
              const float sg = rbf(gg / (1.0f + expf(-gg)));
              st_unit(out + (row >> 1), make_unit(__fmul_rn(sg, uu), tag));
            }
          }
          t_fold += clock64() - m2;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < DA_M_PT; ++i) gen_base[i] += (nt + DA_M_PT - 1 - i) / DA_M_PT;
    cbar();
    total += clock64() - c0;
  }
  if (tid == 0) { cyc[0] = total / reps; cyc[1] = t_mma / reps; cyc[2] = t_sig / reps; cyc[3] = t_fold / reps; }
}
int main() {
  uint32_t *out; long long *cyc; cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { const char *name; int rows, K, epi; } cases[] = {{"qkv (28 x 1024, store)", 28, 1024, ME_STORE}, {"wo (7 x 2048, residual)", 7, 2048, ME_RESIDUAL},
                                                             {"w13 (42 x 1024, swiglu)", 42, 1024, ME_SWIGLU}, {"w2 (7 x 3072, residual)", 7, 3072, ME_RESIDUAL},
                                                             {"fast qkv (14 x 1024)", 14, 1024, ME_STORE}};
  for (auto &c : cases) {
    k<<<1, DA_M_THREADS, 200 * 1024>>>(c.rows, c.K, c.epi, 200, out, cyc);
    long long h[4]; cudaError_t e = cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("%-26s: %5lld cycles per phase (barrier to barrier); warp 0: mma %lld, signal %lld, fold %lld  %s\n", c.name, h[0], h[1], h[2], h[3], e ? cudaGetErrorString(e) : "");
  }
  return 0;
}
