// bstep.cuh -- the batched decode step as a PERSISTENT kernel: one CTA per SM walks a table of phases.
//
// The per-kernel graph of batch_host.cuh launches 536 dependent kernels per step (s1-mini); at 32 requests each of them moves a few
// MB at most, so the step is bound by the ~8 us a dependent kernel costs (launch, set-up, first TMA round trip, drain), not by
// bytes or flops (profiles/r02_ncu_batch_decode.txt).  Here the SAME device code (the b_*_body functions of batch.cuh, the tcgen05
// main loop and tc_epilogue of gemm_tc.cuh) runs inside one cooperative launch:
//
//   * a phase = what used to be one kernel; its (virtual) grid is dealt round-robin to the CTAs; a grid barrier (one atomic + an
//     acquire spin per CTA) separates phases;
//   * warps 0-7 (256 threads, named barrier 1) run the phase bodies; in GEMM phases warps 0-3 are the epilogue (TMEM lane quarters),
//     lane 0 of warp 4 issues tcgen05.mma;
//   * warp 8 is the TMA producer and is NOT part of the phase barrier: while the compute warps are still in phase p it already streams
//     the WEIGHT tiles of the next GEMM phase into the ring (weights do not depend on earlier phases) and adds the activation tiles as
//     soon as the barrier in front of that phase has been passed ("go").  The first HBM round trip of every GEMM is therefore hidden;
//   * TMEM (BN columns), the mbarrier ring and the tensor maps (an array in global memory) are set up once per launch.
//
// The arithmetic, the split-K partition and the summation orders are those of the per-kernel path: the two produce identical bits
// (tests/test_gpu_batch.py).  Every wait is bounded and a timed-out wait aborts the whole grid through the fault flag.
#pragma once
#include "batch.cuh"
#include "gemm_tc.cuh"

namespace da {

enum { BP_EMBED = 0, BP_NORM, BP_QKV_POST, BP_ATTN, BP_FAST_ATTN, BP_FAST_SAMPLE, BP_GEMM };

#define DA_BS_THREADS 288
#define DA_BS_COMPUTE 256
#define DA_BS_WAIT_CYCLES 400000000ll      // ~0.2 s: a lost hand-over raises the fault flag instead of hanging the GPU
typedef BlockNamed<1, DA_BS_COMPUTE> BsBlock;

struct BPhase {
  int kind, gx, gy, gz;        // virtual grid of the phase (GEMM: row tiles, column tiles, K splits)
  int map_w, map_x, pad0, pad1;      // GEMM: indices into the tensor-map array
  union { BEmbedArgs embed; BNormArgs norm; BQkvPostArgs post; BAttnArgs attn; BFastAttnArgs fattn; BFastSampleArgs fsample; GemmTcArgs gemm; } u;
};

struct BStepArgs {
  const BPhase *phases; int n_phases;
  const CUtensorMap *maps;
  unsigned int *gbar;          // grid barrier counter, zero at launch
  int *err;
  int stages;                  // ring depth
  int no_prefetch;             // 1: the producer waits for the phase barrier before it touches the ring (experiment switch)
  long long *tl;               // optional: per phase, CTA 0's clock64 at phase start and at the end of its own work (profiling)
};

__device__ __forceinline__ bool bs_wait(uint64_t *bar, uint32_t parity, volatile int *abortp) {
  if (*abortp) return false;
  const long long t0 = clock64();
  for (;;) {
    uint32_t done = 0;
    for (int it = 0; it < 256 && !done; ++it) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
    if (done) return true;
    if (*abortp) return false;
    if (clock64() - t0 > DA_BS_WAIT_CYCLES) { *abortp = 1; return false; }
  }
}

static inline size_t bstep_smem(int BN, int stages, size_t body_bytes) {
  const size_t stg = (size_t)BN * DA_TC_BM * 4;
  return 1024 + (size_t)stages * (DA_TC_A_BYTES + (size_t)BN * 128) + (body_bytes > stg ? body_bytes : stg);
}

template <int BN>
__global__ void __launch_bounds__(DA_BS_THREADS, 1) bstep_kernel(const BStepArgs k) {
  static_assert(BN == 32 || BN == 64 || BN == 128, "TMEM allocations are powers of two >= 32 columns");
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[DA_TC_MAX_STAGES], empty_bar[DA_TC_MAX_STAGES], accum_bar, accfree_bar, attn_bar[DA_B_NBUF];
  __shared__ uint32_t s_tmem, s_last;
  __shared__ int s_abort, s_go;
  __shared__ __align__(16) BPhase s_ph[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ST = k.stages, G = gridDim.x, cta = blockIdx.x;
  unsigned char *sbase = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *sA = sbase, *sB = sbase + (size_t)ST * DA_TC_A_BYTES;
  unsigned char *body = sB + (size_t)ST * BN * 128;      // phase bodies' scratch; doubles as the epilogue's fp32 staging tile
  constexpr uint32_t STAGE_BYTES = DA_TC_A_BYTES + BN * 128;
  constexpr int PH_WORDS = (int)(sizeof(BPhase) / 4);
  volatile int *abortp = &s_abort, *gop = &s_go;

  if (tid == 0) {
    for (int i = 0; i < ST; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&accum_bar, 1); mbar_init(&accfree_bar, 128); for (int i = 0; i < DA_B_NBUF; ++i) mbar_init(&attn_bar[i], 1);
    s_abort = 0; s_go = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(&s_tmem, BN);
  for (int i = tid; i < PH_WORDS; i += DA_BS_THREADS) reinterpret_cast<uint32_t *>(&s_ph[0])[i] = reinterpret_cast<const uint32_t *>(k.phases)[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp == 8) {
    // ===== TMA producer: runs ahead of the phase barrier with the weight tiles of the next GEMM phase =====
    if (lane == 0) {
      uint64_t pol_keep, pol_first;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
      uint32_t it = 0;
      bool ok = true;
      for (int p = 0; p < k.n_phases && ok; ++p) {
        const BPhase *gp = k.phases + p;
        if (__ldg(&gp->kind) != BP_GEMM) continue;
        const int gx = __ldg(&gp->gx), gy = __ldg(&gp->gy), nz = __ldg(&gp->gz), n_units = gx * gy * nz;
        const CUtensorMap *mw = k.maps + __ldg(&gp->map_w), *mx = k.maps + __ldg(&gp->map_x);
        const int nkb_all = __ldg(&gp->u.gemm.K) / DA_TC_BK;
        const uint64_t pol_w = __ldg(&gp->u.gemm.w_keep) ? pol_keep : pol_first;
        int npend = 0, pslot[DA_TC_MAX_STAGES], pc0[DA_TC_MAX_STAGES], pc1[DA_TC_MAX_STAGES];      // stages whose activation tile waits for "go"
        bool gone = *gop >= p;
        auto wait_go = [&]() {
          const long long t0 = clock64();
          while (*gop < p) {
            if (*abortp) { ok = false; return; }
            if (clock64() - t0 > DA_BS_WAIT_CYCLES) { *abortp = 1; atomicExch(k.err, 9); ok = false; return; }
            __nanosleep(20);
          }
          __threadfence();
          asm volatile("fence.proxy.async;" ::: "memory");      // other CTAs' generic-proxy writes -> this thread's async-proxy (TMA) reads
        };
        auto release = [&]() {
          wait_go(); if (!ok) return;
          for (int j = 0; j < npend; ++j) tma_load_2d(sB + (size_t)pslot[j] * BN * 128, mx, pc0[j], pc1[j], &full_bar[pslot[j]], pol_keep);
          npend = 0; gone = true;
        };
        if (k.no_prefetch && !gone) { wait_go(); gone = true; }
        for (int u = cta; u < n_units && ok; u += G) {
          const int tile_m = u % gx, tile_n = (u / gx) % gy, z = u / (gx * gy);
          const int kb0 = (nkb_all * z) / nz, kb1 = (nkb_all * (z + 1)) / nz;
          for (int i = kb0; i < kb1 && ok; ++i, ++it) {
            if (!gone && npend == ST) { release(); if (!ok) break; }      // the ring is full of half-issued stages
            const int s = (int)(it % (uint32_t)ST); const uint32_t ph = (it / (uint32_t)ST) & 1u;
            ok = bs_wait(&empty_bar[s], ph ^ 1u, abortp);
            if (!ok) { atomicExch(k.err, 5); k.err[2] = p; break; }
            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
            tma_load_2d(sA + (size_t)s * DA_TC_A_BYTES, mw, i * DA_TC_BK, tile_m * DA_TC_BM, &full_bar[s], pol_w);
            if (gone) tma_load_2d(sB + (size_t)s * BN * 128, mx, i * DA_TC_BK, tile_n * BN, &full_bar[s], pol_keep);
            else { pslot[npend] = s; pc0[npend] = i * DA_TC_BK; pc1[npend] = tile_n * BN; ++npend; }
          }
        }
        if (ok && !gone) release();
      }
    }
  } else {
    // ===== compute warps =====
    uint32_t it = 0, uc = 0;      // running k-block / unit counters: the mbarrier parities continue across units and phases
    uint32_t attn_phase[DA_B_NBUF] = {};
    for (int p = 0; p < k.n_phases; ++p) {
      const BPhase &ph = s_ph[p & 1];
      if (k.tl && cta == 0 && tid == 0) k.tl[2 * p] = clock64();
      if (!*abortp) {
        const int gx = ph.gx, gy = ph.gy, n_units = gx * gy * ph.gz;
        if (ph.kind == BP_GEMM) {
          const GemmTcArgs &a = ph.u.gemm;
          const int nz = ph.gz, nkb_all = a.K / DA_TC_BK;
          for (int u = cta; u < n_units; u += G, ++uc) {
            const int tile_m = u % gx, tile_n = (u / gx) % gy, z = u / (gx * gy);
            const int kb0 = (nkb_all * z) / nz, kb1 = (nkb_all * (z + 1)) / nz, nk = kb1 - kb0;
            if (warp == 4) {
              if (lane == 0) {
                const uint32_t idesc = umma_idesc_bf16(BN);
                bool ok = bs_wait(&accfree_bar, (uc & 1u) ^ 1u, abortp);      // the epilogue has drained the previous unit's accumulator
                tc_fence_after();
                for (int i = 0; i < nk && ok; ++i, ++it) {
                  const int s = (int)(it % (uint32_t)ST); const uint32_t par = (it / (uint32_t)ST) & 1u;
                  ok = bs_wait(&full_bar[s], par, abortp);
                  tc_fence_after();
                  const uint64_t da0 = umma_desc_sw128(smem_u32(sA + (size_t)s * DA_TC_A_BYTES));
                  const uint64_t db0 = umma_desc_sw128(smem_u32(sB + (size_t)s * BN * 128));
#pragma unroll
                  for (int kk = 0; kk < DA_TC_BK / 16; ++kk) umma_bf16(tmem, da0 + (uint64_t)(kk * 2), db0 + (uint64_t)(kk * 2), idesc, (uint32_t)((i | kk) != 0));
                  umma_commit(&empty_bar[s]);
                }
                umma_commit(&accum_bar);
                if (!ok) { atomicExch(k.err, 6); k.err[1] = p; }
              }
              __syncwarp();
            } else if (warp < 4) {
              const bool ok = bs_wait(&accum_bar, uc & 1u, abortp);
              tc_fence_after();
              if (!ok) { atomicExch(k.err, 7); k.err[1] = p; }
              tc_epilogue<BN>(a, tmem, reinterpret_cast<float *>(body), tile_m, tile_n, z, nz, tile_n * gx + tile_m, tid, lane, warp, &s_last, ok, &accfree_bar);
            }
          }
        } else {
          for (int u = cta; u < n_units; u += G) {
            const int bx = u % gx, by = (u / gx) % gy, bz = u / (gx * gy);
            switch (ph.kind) {
              case BP_EMBED: b_embed_body<BsBlock>(ph.u.embed, bx, by, bz, gy, body); break;
              case BP_NORM: b_rmsnorm_body<BsBlock>(ph.u.norm, bx, by, bz, gy, body); break;
              case BP_QKV_POST: b_qkv_post_body<BsBlock>(ph.u.post, bx, by, bz, gy, body); break;
              case BP_ATTN:
                if (ph.u.attn.use_mma && ph.u.attn.hd == 128) b_attn_body<BsBlock, true, 128>(ph.u.attn, bx, by, bz, gy, body, attn_bar, attn_phase, false, k.maps + ph.map_w, k.maps + ph.map_x);
                else if (ph.u.attn.use_mma) b_attn_body<BsBlock, true, 64>(ph.u.attn, bx, by, bz, gy, body, attn_bar, attn_phase, false, k.maps + ph.map_w, k.maps + ph.map_x);
                else b_attn_body<BsBlock>(ph.u.attn, bx, by, bz, gy, body, attn_bar, attn_phase, false);
                break;
              case BP_FAST_ATTN: b_fast_attn_body<BsBlock>(ph.u.fattn, bx, by, bz, gy, body); break;
              case BP_FAST_SAMPLE: b_fast_sample_body<BsBlock>(ph.u.fsample, bx, by, bz, gy, body); break;
              default: break;
            }
            if (u + G < n_units) BsBlock::sync();      // the next unit reuses the scratch
          }
        }
      }
      if (k.tl && cta == 0 && tid == 0) k.tl[2 * p + 1] = clock64();
      if (p + 1 == k.n_phases) break;
      // ---- grid barrier; the next phase's descriptor is fetched under it ----
      asm volatile("fence.proxy.async;" ::: "memory");      // this phase's generic-proxy writes -> later TMA reads of them (any CTA)
      BsBlock::sync();
      if (tid == 0) {
        __threadfence();
        atomicAdd(k.gbar, 1u);
        const unsigned target = (unsigned)(p + 1) * (unsigned)G;
        if (!*abortp) {
          const long long t0 = clock64();
          int polls = 0;
          for (;;) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(k.gbar) : "memory");
            if (v >= target) break;
            if ((++polls & 63) == 0) {
              if (*reinterpret_cast<volatile int *>(k.err) != 0) { *abortp = 1; break; }
              if (clock64() - t0 > DA_BS_WAIT_CYCLES) { *abortp = 1; atomicExch(k.err, 8); k.err[1] = p; break; }
            }
          }
        }
        __threadfence();
        *gop = p + 1;
      } else if (warp == 1) {
        for (int i = lane; i < PH_WORDS; i += 32)
          reinterpret_cast<uint32_t *>(&s_ph[(p + 1) & 1])[i] = __ldcg(reinterpret_cast<const uint32_t *>(k.phases + (p + 1)) + i);
      }
      BsBlock::sync();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, BN);
}

}  // namespace da
