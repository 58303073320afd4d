// gemm_tc.cuh -- the dense contractions of the path on 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Where the decode path really is a GEMM -- the prefill of a T-position prompt (inference.py:353-362 runs the whole prompt
// through forward_generate in ONE call) and batched decode (B requests advance together) -- every nn.Linear of the model
// (llama.py:190, 240, 283, 449, 577) becomes
//
//        Y[n][r] = sum_k W[r][k] * X[n][k]            W: [rows][K] weights (row-major, K contiguous, exactly as stored)
//                                                     X: [ncols][K] activations: one row per request / prompt position
//
// computed "swap-AB": the WEIGHT tile is the M = 128 operand of tcgen05.mma (streamed once from HBM through a TMA ring with
// the 128-byte swizzle), the activations are the N operand (N = 32 ... 256 columns, re-read from L2 by every CTA), the fp32
// accumulator lives in TMEM (128 lanes x N columns) and is read back with tcgen05.ld by four epilogue warps, one row per
// thread.  bs = 32 decode therefore streams each weight byte ONCE per step for all 32 requests.
//
//   warp 0     producer: one elected lane arms the stage's mbarrier and issues the 2-D TMA loads (cp.async.bulk.tensor).  The
//              weight tiles of the first `stages` k-blocks are requested BEFORE griddepcontrol.wait: weights do not depend on
//              the previous kernel, so the stream of kernel N+1 starts under the tail of kernel N (programmatic dependent launch)
//   warp 1     allocates TMEM; one elected lane issues tcgen05.mma (4 per 64-element k-block) and commits to the stage's
//              `empty` barrier (releases the smem slot) and finally to the accumulator barrier
//   warps 2-5  epilogue: TMEM -> registers -> shared memory (fp32, [column][row]) -> 8 rows per thread -> (bias, residual, SwiGLU) ->
//              bf16 -> 16-byte stores, laid out [n][row] so that the output is directly the X operand of the next GEMM
//
// Split-K (gridDim.z > 1) for the matrices with few row tiles (wo, w2: 8 tiles): every split writes its fp32 partial tile to
// a workspace, the LAST split to arrive (ticket) adds the partials IN SPLIT ORDER and runs the epilogue -- deterministic.
//
// Rounding points are the reference's: the linear's output is rounded to bf16 (+ bias in fp32 before the
// rounding, like cuBLAS's epilogue), the residual add and SwiGLU round again.  The ORDER of the fp32 sum over k is the tensor
// core's; it is the same for every column, so a request's result does not depend on what else is in the batch.
// Every wait is bounded: a lost transaction raises the device fault flag instead of hanging the GPU.
#pragma once
#include <cuda.h>

#include "attention.cuh"
#include "common.cuh"

namespace da {

enum { TE_STORE = 0, TE_RESIDUAL = 1, TE_SWIGLU = 2 };

#define DA_TC_THREADS 192
#define DA_TC_BM 128                      // weight rows per tile = UMMA M
#define DA_TC_BK 64                       // k-block: 64 bf16 = one 128-byte swizzle row
#define DA_TC_A_BYTES (DA_TC_BM * 128)    // 16 KB per stage
#define DA_TC_MAX_STAGES 12
#ifndef DA_TC_MIN_BLOCKS
#define DA_TC_MIN_BLOCKS 1             // co-resident GEMM CTAs per SM the register allocation must allow (166 registers at 1: two fit)
#endif

struct GemmTcArgs {
  int rows, K, ncols;        // weight rows, contraction length, valid activation rows
  int stages;                // smem ring depth
  int epi;                   // TE_*
  int ld_out;                // elements between consecutive activation rows of out / res (rows, or rows / 2 for SwiGLU)
  int w_keep;                // 1: the weight is re-read soon (fast stack): L2 evict_last; 0: streamed once: evict_first
  const bf16 *bias;          // [rows] or null
  const bf16 *res;           // TE_RESIDUAL: [ncols][ld_out]
  bf16 *out;                 // [ncols][ld_out]
  const bf16 *xraw;          // XN kernels: un-normalised activations [ncols][K]; the kernel applies the custom RMSNorm itself
  const bf16 *norm_w;        //             norm weight [K]
  float eps;
  const float *ssq_in;       // XN kernels, optional: per column ssq_n partial sums of squares of xraw's row (written by the producing GEMM)
  int ssq_n;
  float *ssq_out;            // optional: this GEMM's epilogue writes, per column and row tile, the sum of squares of the bf16 outputs: [col][ssq_ld]
  int ssq_ld;
  float *ws;                 // split-K partials [tile][split][BN][128]
  unsigned int *tickets;     // split-K arrival counters, one per tile, zero between launches
  int *err;                  // device fault flag (DAState::err or a stand-alone word)
  long long *dbg;            // optional clock64 stamps, 8 per CTA (tests/cuda/gemm_tc_check.cu): start, set-up done, dependency wait over,
                             // first stage landed, last MMA issued, accumulator complete, epilogue done
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major; issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                 "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                 "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// shared-memory matrix descriptor of a K-major tile in the 128-byte-swizzle layout TMA writes: rows of 128 bytes, 8-row groups
// 1024 bytes apart (SBO), version 1 (Blackwell), layout SWIZZLE_128B.  Advancing by 16 elements along K = +32 bytes on the start
// address (the hardware applies the XOR swizzle to the absolute address; tiles are 1024-byte aligned).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor: D fp32, A and B bf16, both K-major, M = 128, N = n
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(DA_TC_BM >> 4) << 24);
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  for (int it = 0; it < (1 << 20) && !done; ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
  return done != 0;
}

// the ring doubles as the epilogue's fp32 staging tile ([BN][128] floats)
static inline size_t gemm_tc_smem(int BN, int stages) {
  const size_t ring = (size_t)stages * (DA_TC_A_BYTES + (size_t)BN * 128), stg = (size_t)BN * DA_TC_BM * 4;
  return (ring > stg ? ring : stg) + 1024;
}
// CL kernels: the ring, then the receive buffer of the cluster reduction (BN x 128 floats however many splits)
static inline size_t gemm_tc_smem_cl(int BN, int stages) { return (size_t)stages * (DA_TC_A_BYTES + (size_t)BN * 128) + (size_t)BN * DA_TC_BM * 4 + 1024; }
// XN kernels: a ring of weight tiles only, then the resident activation operand of the CTA's nk k-blocks
static inline size_t gemm_tc_smem_xn(int BN, int stages, int nk) {
  const size_t ring = (size_t)stages * DA_TC_A_BYTES, stg = (size_t)BN * DA_TC_BM * 4;
  return (ring > stg ? ring : stg) + (size_t)nk * BN * 128 + 1024;
}

// bias / residual / SwiGLU on the 8 rows [row0, row0 + 8) of column n_base + n_l, 16 contiguous bytes out (shared by all epilogues)
__device__ __forceinline__ void tc_finish(const GemmTcArgs &a, const float *acc, int tile_m, int n_base, int n_l, int g8, int lane) {
        const int row0 = tile_m * DA_TC_BM + g8;
        if (row0 >= a.rows) return;                                  // rows is a multiple of 8
        float y[8];
        if (a.bias) {
          float bb[8]; unpack8(*reinterpret_cast<const uint4 *>(a.bias + row0), bb);
#pragma unroll
          for (int i2 = 0; i2 < 8; ++i2) y[i2] = rbf(acc[i2] + bb[i2]);
        } else {
#pragma unroll
          for (int i2 = 0; i2 < 8; ++i2) y[i2] = rbf(acc[i2]);       // the linear's bf16 output
        }
        const size_t n = (size_t)(n_base + n_l);
        if (a.epi == TE_SWIGLU) {
          // rows come interleaved (2j: w1 = gate, 2j+1: w3 = up; engine.cu plan_layer): bf16(bf16(silu(g)) * u)   llama.py:190
          float o[4];
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2) { const float gg = y[2 * i2]; o[i2] = __fmul_rn(rbf(gg / (1.0f + expf(-gg))), y[2 * i2 + 1]); }
          uint2 u; u.x = (uint32_t)f2bits(o[0]) | ((uint32_t)f2bits(o[1]) << 16); u.y = (uint32_t)f2bits(o[2]) | ((uint32_t)f2bits(o[3]) << 16);
          *reinterpret_cast<uint2 *>(a.out + n * a.ld_out + (row0 >> 1)) = u;
        } else {
          if (a.epi == TE_RESIDUAL) {                                // llama.py:329-330
            float rr[8]; unpack8(*reinterpret_cast<const uint4 *>(a.res + n * a.ld_out + row0), rr);
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2) y[i2] = rr[i2] + y[i2];
          }
          uint4 u;
          u.x = (uint32_t)f2bits(y[0]) | ((uint32_t)f2bits(y[1]) << 16); u.y = (uint32_t)f2bits(y[2]) | ((uint32_t)f2bits(y[3]) << 16);
          u.z = (uint32_t)f2bits(y[4]) | ((uint32_t)f2bits(y[5]) << 16); u.w = (uint32_t)f2bits(y[6]) | ((uint32_t)f2bits(y[7]) << 16);
          *reinterpret_cast<uint4 *>(a.out + n * a.ld_out + row0) = u;
          if (a.ssq_out) {
            // statistics for the RMSNorm that consumes this output (llama.py:172-177 squares the bf16 values in fp32): the 16 items of a
            // column in this 128-row tile sit in the 16 lanes of a half-warp -- fixed-order butterfly, one partial per (column, row tile)
            float ss = 0.f;
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2) { const float r = rbf(y[i2]); ss = fmaf(r, r, ss); }
            const unsigned hm = 0xFFFFu << (lane & 16);
            ss += __shfl_xor_sync(hm, ss, 8); ss += __shfl_xor_sync(hm, ss, 4); ss += __shfl_xor_sync(hm, ss, 2); ss += __shfl_xor_sync(hm, ss, 1);
            if ((lane & 15) == 0) a.ssq_out[n * a.ssq_ld + tile_m] = ss;
          }
        }
}

// ---- the epilogue, shared by the stand-alone kernel and the persistent batched-step kernel (bstep.cuh) ---------------------------------
// Runs on the four warps that own the TMEM lane quarters (128 threads; te = 0..127 their index, wq = warp % 4), after the accumulator
// barrier has fired.  `stg` = 16-byte aligned shared memory for BN x 128 floats; named barrier 2 is the epilogue group's.
//   (A) the fp32 tile leaves TMEM as [column][row]: into the split-K workspace (nz > 1) or into `stg`;
//   (B) a compact loop over (column, group of 8 rows) items reads it back 8 rows at a time, applies bias / residual / SwiGLU and writes
//       16 contiguous bytes per item.  Straight-line code that runs once per CTA executes at instruction-fetch speed (measured: the
//       32-column unrolled store sequence cost 7,000 cycles per chunk), so (B) is a LOOP on purpose.
template <int BN>
__device__ __forceinline__ void tc_epilogue(const GemmTcArgs &a, uint32_t tmem, float *stg, int tile_m, int tile_n, int z, int nz, int ntile_lin,
                                            int te, int lane, int wq, uint32_t *s_last_p, bool ok_in, uint64_t *accfree = nullptr) {
  const int r_in = wq * 32 + lane;
  const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16);
  uint32_t &s_last = *s_last_p;
  // the waits in front of the epilogue are bounded per thread: agree on the outcome, the barriers below need all 128 threads
  uint32_t ok_all;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.and.pred q, 2, 128, p;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok_all) : "r"((uint32_t)ok_in) : "memory");
  const bool ok = ok_all != 0u;
    const int n_base = tile_n * BN, ncols_here = min(BN, a.ncols - n_base);
    bool is_final = true;
    if (ok) {
      float *dstp = nz > 1 ? a.ws + ((size_t)ntile_lin * nz + z) * BN * DA_TC_BM : stg;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (c0 >= ncols_here) break;
        uint32_t v[32]; tmem_ld32(taddr + (uint32_t)c0, v);
        if (nz > 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) __stcg(dstp + (size_t)(c0 + j) * DA_TC_BM + r_in, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) dstp[(c0 + j) * DA_TC_BM + r_in] = __uint_as_float(v[j]);
        }
      }
    }
    if (accfree) { tc_fence_before(); mbar_arrive(accfree); }      // persistent kernel: the accumulator may be overwritten by the next unit
    if (ok) {
      if (nz > 1) {
        __threadfence();
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (te == 0) s_last = (atomicAdd(a.tickets + ntile_lin, 1u) == (unsigned)nz - 1) ? 1u : 0u;
        asm volatile("bar.sync 2, 128;" ::: "memory");
        is_final = s_last != 0u;
        if (is_final) { __threadfence(); if (te == 0) a.tickets[ntile_lin] = 0u; }
      } else asm volatile("bar.sync 2, 128;" ::: "memory");
    }
    if (is_final && ok) {
      const float *wsp = a.ws + (size_t)ntile_lin * nz * BN * DA_TC_BM;
      const int n_items = ncols_here * (DA_TC_BM / 8);
      auto finish = [&](const float *acc, int n_l, int g8) { tc_finish(a, acc, tile_m, n_base, n_l, g8, lane); };
      if (nz > 1) {
        // split-K: the partials of TWO items (2 x nz x 32 bytes, nz <= 8) are requested before the first add, so the reduction costs one
        // L2 round trip per pair of items instead of one per partial; they are added in split order (deterministic)
#pragma unroll 1
        for (int item = te; item < n_items; item += 256) {
          const int item1 = item + 128; const bool two = item1 < n_items;
          const int nl0 = item >> 4, g0 = (item & 15) * 8, nl1 = two ? item1 >> 4 : nl0, g1 = two ? (item1 & 15) * 8 : g0;
          float4 p[2][8][2];
#pragma unroll
          for (int zz = 0; zz < 8; ++zz) {
            if (zz < nz) {
              const float4 *q0 = reinterpret_cast<const float4 *>(wsp + ((size_t)zz * BN + nl0) * DA_TC_BM + g0);
              const float4 *q1 = reinterpret_cast<const float4 *>(wsp + ((size_t)zz * BN + nl1) * DA_TC_BM + g1);
              p[0][zz][0] = __ldcg(q0); p[0][zz][1] = __ldcg(q0 + 1); p[1][zz][0] = __ldcg(q1); p[1][zz][1] = __ldcg(q1 + 1);
            }
          }
#pragma unroll
          for (int w2 = 0; w2 < 2; ++w2) {
            if (w2 == 1 && !two) break;
            float acc[8];
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2) acc[i2] = 0.f;
#pragma unroll
            for (int zz = 0; zz < 8; ++zz) {
              if (zz < nz) {
                acc[0] += p[w2][zz][0].x; acc[1] += p[w2][zz][0].y; acc[2] += p[w2][zz][0].z; acc[3] += p[w2][zz][0].w;
                acc[4] += p[w2][zz][1].x; acc[5] += p[w2][zz][1].y; acc[6] += p[w2][zz][1].z; acc[7] += p[w2][zz][1].w;
              }
            }
            finish(acc, w2 ? nl1 : nl0, w2 ? g1 : g0);
          }
        }
      } else {
#pragma unroll 1
        for (int item = te; item < n_items; item += 128) {
          const int n_l = item >> 4, g8 = (item & 15) * 8;
          const float4 *pz = reinterpret_cast<const float4 *>(stg + n_l * DA_TC_BM + g8);
          const float4 p0 = pz[0], p1 = pz[1];
          const float acc[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
          finish(acc, n_l, g8);
        }
      }
    }
}

// grid (row tiles, column tiles, K splits)
// XN = true (BN = 32, batched decode): the activation operand is NOT loaded by TMA; the epilogue warps, idle during the main loop,
// read the un-normalised rows, apply the reference's RMSNorm (llama.py:172-177: fp32 normalise, round, x weight, round -- every CTA
// recomputes the row statistics, 64 KB of L2 reads) and write the CTA's k-range into shared memory in the same 128-byte-swizzle
// K-major layout TMA would have produced.  That removes one kernel (and one kernel boundary, ~5 us in a dependent chain) per norm.
template <int BN, bool XN = false, bool CL = false>
__global__ void __launch_bounds__(DA_TC_THREADS, XN ? 2 : DA_TC_MIN_BLOCKS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const GemmTcArgs a) {
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "TMEM allocations are powers of two >= 32 columns");
  static_assert(!XN || BN == 32, "the fused-norm operand staging deals 8 rows to each of the four epilogue warps");
  static_assert(!(XN && CL), "the cluster reduction is for the plain TMA-fed GEMM");
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[DA_TC_MAX_STAGES], empty_bar[DA_TC_MAX_STAGES], accum_bar, xready_bar;
  __shared__ uint32_t s_tmem, s_last;
  __shared__ float s_inv[XN ? BN : 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y, z = blockIdx.z, nz = gridDim.z;
  const int stages = a.stages;
  unsigned char *sbase = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *sA = sbase, *sB = sbase + (size_t)stages * DA_TC_A_BYTES;
  const int nkb_all = a.K / DA_TC_BK;
  const int kb0 = (nkb_all * z) / nz, kb1 = (nkb_all * (z + 1)) / nz, nk = kb1 - kb0;
  constexpr uint32_t STAGE_BYTES = DA_TC_A_BYTES + (XN ? 0 : BN * 128);      // XN: sB holds all k-blocks of this CTA, written by the epilogue warps

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  long long *dbg = a.dbg ? a.dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (dbg && tid == 0) dbg[0] = clock64();
  if (tid == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&accum_bar, 1);
    if (XN) mbar_init(&xready_bar, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmX);
  }
  if (warp == 1) tmem_alloc(&s_tmem, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  bool ok = true;
  if (dbg && tid == 0) dbg[1] = clock64();
  // CL: the K splits of a tile form a thread-block cluster (cluster dims (1, 1, gridDim.z)) and reduce through distributed shared memory
  // instead of a global workspace + ticket.  Barrier 1 (arrive here, wait before the first remote store) only proves that every CTA of
  // the cluster has started; barrier 2 publishes the partials.  All 192 threads take part in both.
  if (CL) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      uint64_t pol_w, pol_x;
      if (a.w_keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_w));
      else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_w));
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_x));
      const int npre = nk < stages ? nk : stages;
      for (int i = 0; i < npre; ++i) {
        mbar_expect_tx(&full_bar[i], STAGE_BYTES);
        tma_load_2d(sA + (size_t)i * DA_TC_A_BYTES, &tmW, (kb0 + i) * DA_TC_BK, tile_m * DA_TC_BM, &full_bar[i], pol_w);
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");
      if (dbg) dbg[2] = clock64();
      if (!XN) for (int i = 0; i < npre; ++i) tma_load_2d(sB + (size_t)i * BN * 128, &tmX, (kb0 + i) * DA_TC_BK, tile_n * BN, &full_bar[i], pol_x);
      for (int i = npre; i < nk && ok; ++i) {
        const int s = i % stages; const uint32_t ph = (uint32_t)(i / stages) & 1u;
        ok = mbar_wait_bounded(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_2d(sA + (size_t)s * DA_TC_A_BYTES, &tmW, (kb0 + i) * DA_TC_BK, tile_m * DA_TC_BM, &full_bar[s], pol_w);
        if (!XN) tma_load_2d(sB + (size_t)s * BN * 128, &tmX, (kb0 + i) * DA_TC_BK, tile_n * BN, &full_bar[s], pol_x);
      }
      if (!ok) atomicExch(a.err, 5);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BN);
      if (XN) { ok = mbar_wait_bounded(&xready_bar, 0u); tc_fence_after(); }
      for (int i = 0; i < nk && ok; ++i) {
        const int s = i % stages; const uint32_t ph = (uint32_t)(i / stages) & 1u;
        ok = mbar_wait_bounded(&full_bar[s], ph);
        tc_fence_after();
        if (dbg && i == 0) dbg[3] = clock64();
        const uint64_t da0 = umma_desc_sw128(smem_u32(sA + (size_t)s * DA_TC_A_BYTES));
        const uint64_t db0 = umma_desc_sw128(smem_u32(sB + (size_t)(XN ? i : s) * BN * 128));
#pragma unroll
        for (int k = 0; k < DA_TC_BK / 16; ++k) umma_bf16(tmem, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc, (uint32_t)((i | k) != 0));
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&accum_bar);
      if (dbg) dbg[4] = clock64();
      if (!ok) atomicExch(a.err, 6);
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 2..5 own TMEM lanes 32 * (warp % 4) .. + 31 =====
    // (A) the fp32 tile leaves TMEM as [column][row]: into the split-K workspace (gridDim.z > 1) or into the shared memory
    //     of the ring, which is idle once the accumulator barrier has fired;
    // (B) a compact loop over (column, group of 8 rows) items reads it back 8 rows at a time, applies bias / residual / SwiGLU
    //     and writes 16 contiguous bytes per item.  Straight-line code that runs once per CTA executes at instruction-fetch
    //     speed (measured: the 32-column unrolled store sequence cost 7,000 cycles per chunk), so (B) is a LOOP on purpose.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int te = tid - 64, wq = warp & 3, r_in = wq * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16);
    if (XN) {
      // ---- (X) the normalised activation operand of this CTA's k-range, built in shared memory ---------------------------------------
      // row statistics: warp wq owns rows wq, wq + 4, ...; per lane the chunks lane, lane + 32, ... in ascending order, then a butterfly
      // (the summation order of b_rmsnorm_kernel, so fused and unfused paths agree bit for bit)
      const int K = a.K, cpl = K >> 8;      // 16-byte chunks per lane
      if (a.ssq_in) {
        // the producing GEMM's epilogue left ssq_n partial sums of squares per column: add them in tile order
        if (te < BN) {
          float t = 0.f;
          if (te < a.ncols) for (int j = 0; j < a.ssq_n; ++j) t += a.ssq_in[(size_t)te * a.ssq_n + j];
          s_inv[te] = rsqrtf(t * (1.0f / (float)K) + a.eps);
        }
      } else {
      float ss[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) ss[r] = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < cpl; c0 += 4) {
        uint4 v[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int n = wq + 4 * r;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n < a.ncols && c0 + j < cpl) v[r][j] = *reinterpret_cast<const uint4 *>(a.xraw + (size_t)n * K + ((size_t)(c0 + j) * 32 + lane) * 8);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int n = wq + 4 * r;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n < a.ncols && c0 + j < cpl) {
              float f[8]; unpack8(v[r][j], f);
#pragma unroll
              for (int q = 0; q < 8; ++q) ss[r] = fmaf(f[q], f[q], ss[r]);
            }
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float tot = warp_sum(ss[r]);
        if (lane == 0) s_inv[wq + 4 * r] = rsqrtf(tot * (1.0f / (float)K) + a.eps);
      }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // transform + swizzled store: 16-byte chunk c16 of row n of k-block kb goes to  sB + kb * BN * 128 + n * 128 + ((c16 ^ (n & 7)) << 4)
      const int per_row = nk * 8, total = BN * per_row;
      // eight chunks per thread and round: all sixteen 16-byte loads (activations + norm weights) are in flight before the first is
      // used, so a round costs one L2 round trip instead of eight dependent ones
#pragma unroll 1
      for (int c0 = te; c0 < total; c0 += 8 * 128) {
        uint4 xv[8], gv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + j * 128;
          xv[j] = make_uint4(0u, 0u, 0u, 0u); gv[j] = xv[j];
          if (c < total) {
            const int n = c / per_row, kc = c - n * per_row;
            const size_t k0 = (size_t)(kb0 + (kc >> 3)) * DA_TC_BK + (size_t)(kc & 7) * 8;
            if (n < a.ncols) { xv[j] = *reinterpret_cast<const uint4 *>(a.xraw + (size_t)n * K + k0); gv[j] = *reinterpret_cast<const uint4 *>(a.norm_w + k0); }
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + j * 128;
          if (c < total) {
            const int n = c / per_row, kc = c - n * per_row, kb = kc >> 3, c16 = kc & 7;
            uint4 u = make_uint4(0u, 0u, 0u, 0u);
            if (n < a.ncols) {
              float f[8], g[8];
              unpack8(xv[j], f); unpack8(gv[j], g);
              const float inv = s_inv[n];
#pragma unroll
              for (int q = 0; q < 8; ++q) f[q] = rbf(__fmul_rn(rbf(__fmul_rn(f[q], inv)), g[q]));      // .type_as(x), then * weight
              u.x = (uint32_t)f2bits(f[0]) | ((uint32_t)f2bits(f[1]) << 16); u.y = (uint32_t)f2bits(f[2]) | ((uint32_t)f2bits(f[3]) << 16);
              u.z = (uint32_t)f2bits(f[4]) | ((uint32_t)f2bits(f[5]) << 16); u.w = (uint32_t)f2bits(f[6]) | ((uint32_t)f2bits(f[7]) << 16);
            }
            *reinterpret_cast<uint4 *>(sB + (size_t)kb * BN * 128 + (size_t)n * 128 + (size_t)((c16 ^ (n & 7)) << 4)) = u;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core's async proxy
      mbar_arrive(&xready_bar);
    }
    ok = mbar_wait_bounded(&accum_bar, 0u);
    tc_fence_after();
    if (dbg && tid == 64) dbg[5] = clock64();
    if (!ok) atomicExch(a.err, 7);
    if (!CL) tc_epilogue<BN>(a, tmem, reinterpret_cast<float *>(sbase), tile_m, tile_n, z, nz, blockIdx.y * gridDim.x + blockIdx.x, te, lane, wq, &s_last, ok);
    else {
      // ---- reduce-scatter over the cluster: CTA r finishes rows [r * RPC, (r + 1) * RPC) of the tile.  Every thread holds one row of the
      // CTA's partial (its TMEM lane); it stores the row's BN values into the owner's receive buffer [source z][column][row in slice]
      // (remote shared memory; a warp's 32 rows are 32 consecutive floats there).  After barrier 2 every CTA adds its ks slices in split
      // order -- the order of the workspace path, so both give the same bits -- and runs the usual bias / residual / SwiGLU finish.
      const int RPC = DA_TC_BM / nz, n_base = tile_n * BN, ncols_here = min(BN, a.ncols - n_base);
      float *recv = reinterpret_cast<float *>(sbase + (size_t)stages * (DA_TC_A_BYTES + BN * 128));      // [nz][BN][RPC], behind the ring
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      const int r_in = wq * 32 + lane, owner = r_in / RPC, r_sl = r_in - owner * RPC;
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(recv)), "r"(owner));
      const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (c0 >= ncols_here) break;
        uint32_t v[32]; tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote + (uint32_t)((((size_t)z * BN + c0 + j) * RPC + r_sl) * 4)), "r"(v[j]) : "memory");
      }
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      const int groups = RPC / 8, n_items = ncols_here * groups;      // (column, 8 rows of this CTA's slice)
#pragma unroll 1
      for (int item = te; item < n_items; item += 128) {
        const int n_l = item / groups, g = item - n_l * groups;
        float acc[8];
#pragma unroll
        for (int i2 = 0; i2 < 8; ++i2) acc[i2] = 0.f;
#pragma unroll 1
        for (int zz = 0; zz < nz; ++zz) {
          const float4 *pz = reinterpret_cast<const float4 *>(recv + ((size_t)zz * BN + n_l) * RPC + g * 8);
          const float4 p0 = pz[0], p1 = pz[1];
          acc[0] += p0.x; acc[1] += p0.y; acc[2] += p0.z; acc[3] += p0.w; acc[4] += p1.x; acc[5] += p1.y; acc[6] += p1.z; acc[7] += p1.w;
        }
        tc_finish(a, acc, tile_m, n_base, n_l, z * RPC + g * 8, lane);
      }
    }
  }
  if (CL && warp < 2) {      // producer and MMA warps: their share of the two cluster barriers
    __syncwarp();
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (dbg && tid == 64) dbg[6] = clock64();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, BN);
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline PFN_encodeTiled tc_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
  }
  return fn;
}
// [n_rows][K] bf16 row-major matrix, boxes of box_rows x 64 elements with the 128-byte swizzle; false on failure
static inline bool tc_make_map(CUtensorMap *m, const void *base, int64_t n_rows, int64_t K, int box_rows) {
  PFN_encodeTiled fn = tc_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {DA_TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace da
