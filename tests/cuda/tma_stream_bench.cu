// HBM -> shared memory streaming through the bulk-copy engine, as the persistent kernel's producer does it: every CTA pulls
// its contiguous slice of a 319 MB matrix through a ring of `slots` entries; the consumer only waits and releases.
//   mode 0: one bulk copy per 2 KB row, rows land at a padded stride (2K + 16 bytes) -- what mega.cuh does
//   mode 1: one bulk copy per 16-row tile (32 KB contiguous)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__global__ void __launch_bounds__(64, 1) k(const unsigned char *W, long long rows, int K2, int mode, int slots, long long *out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t *full = reinterpret_cast<uint64_t *>(sm), *empty = full + 32;
  unsigned char *ring = sm + 1024;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, bid = blockIdx.x, grid = gridDim.x;
  const uint32_t RS = K2 + 16, slot_bytes = 16 * RS;
  if (tid == 0) { for (int i = 0; i < 32; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const long long r0 = rows * bid / grid, r1 = rows * (bid + 1) / grid;
  const int nt = (int)((r1 - r0 + 15) / 16);
  unsigned long long g0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
  if (w == 0) {          // producer
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (int t = 0; t < nt; ++t) {
      const int s = t % slots; const uint32_t par = (t / slots) & 1;
      if (t >= slots && lane == 0) mbar_wait(&empty[s], par ^ 1);
      __syncwarp();
      const int n = (int)min(16LL, r1 - r0 - 16LL * t);
      if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)n * K2);
      __syncwarp();
      if (mode == 0) { if (lane < n) bulk_g2s(ring + (size_t)s * slot_bytes + lane * RS, W + (r0 + 16LL * t + lane) * K2, K2, &full[s], pol); }
      else if (lane == 0) bulk_g2s(ring + (size_t)s * slot_bytes, W + (r0 + 16LL * t) * K2, (uint32_t)n * K2, &full[s], pol);
    }
  } else {               // consumer: wait, touch, release
    unsigned acc = 0;
    for (int t = 0; t < nt; ++t) {
      const int s = t % slots; const uint32_t par = (t / slots) & 1;
      mbar_wait(&full[s], par);
      acc += reinterpret_cast<const unsigned *>(ring + (size_t)s * slot_bytes)[lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 0x12345678u) out[1] = acc;
  }
  __syncthreads();
  unsigned long long g1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
  if (tid == 0) atomicMax((unsigned long long *)out, g1 - g0);
}
int main() {
  const long long rows = 155776; const int K2 = 2048;
  unsigned char *W; long long *out; cudaMalloc(&W, rows * K2); cudaMemset(W, 1, rows * K2); cudaMalloc(&out, 64);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int mode = 0; mode < 2; ++mode) for (int slots : {2, 3, 5, 6}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(out, 0, 64);
      k<<<sms, 64, 1024 + slots * 16 * (K2 + 16)>>>(W, rows, K2, mode, slots, out);
      long long h = 0; cudaError_t e = cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      if (rep) printf("mode %d (%s) slots %d (%3d KB ring): %.1f us  %.2f TB/s  %s\n", mode, mode ? "one copy per 16-row tile" : "one copy per row, padded", slots, slots * 16 * (K2 + 16) / 1024,
                      h / 1e3, rows * K2 / (h * 1e-9) / 1e12, e ? cudaGetErrorString(e) : "");
    }
  }
  return 0;
}
