// batch.cuh -- the kernels AROUND the tensor-core GEMMs (gemm_tc.cuh) when the decode path runs over many columns at once:
// a column is one REQUEST (batched decode, BASELINE configs[3]) or one PROMPT POSITION (prefill, inference.py:353-362).
//
// Activations are [column][feature] bf16 row-major, which is directly the K-major N operand of the next GEMM.  Every formula
// and rounding point is the one of the batch-1 kernels (attention.cuh, gemv.cuh, misc_kernels.cuh, sampler.cuh): a request's
// result does not depend on what else is in the batch, which is what the parity tests check ("bs = B equals B independent
// bs = 1 oracle runs"; the reference itself is batch 1 only, inference.py:73, 355).
//
//   b_embed_kernel       llama.py:409-429    token + codebook embedding per column
//   b_rmsnorm_kernel     llama.py:172-177    custom RMSNorm (round before the weight multiply), one warp per column
//   b_qkv_post_kernel    llama.py:246-251, 142-149   q/k nn.RMSNorm + RoPE in place, K/V rows into the column's cache (prefill; decode does it in attention)
//   b_attn_kernel        llama.py:258-274    split-KV flash-decode per (column, kv head), fp32 like the math SDPA backend (scalar walk)
//   b_attn_mma_kernel    the same with Q.K^T and P@V on the tensor cores (decode, head_dim 64 / 128)
//   b_fast_attn_kernel   llama.py:285-309    the fast layers' bf16 attention over <= num_codebooks positions
//   b_head_stats_kernel / b_select_kernel / b_fast_sample_kernel    inference.py:30-80, 103-149    penalty, exact nucleus, Exp(1) race
#pragma once
#include "attention.cuh"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "gemv.cuh"
#include "sampler.cuh"

namespace da {

// token (column n, row r) = base[n * sn + r * sr]: decode reads DAState::tok_in of slot n, prefill a prompt column
struct TokSrc { const int *base; long long sn, sr; };
// position of column n = base ? base[n * sn] : pos0 + n
struct PosSrc { const int *base; long long sn; int pos0; };
__device__ __forceinline__ int pos_of(const PosSrc &p, int n) { return p.base ? p.base[(long long)n * p.sn] : p.pos0 + n; }

// ---- embedding ---------------------------------------------------------------------------------------------------------------
struct BEmbedArgs {
  const bf16 *emb, *cb_emb; bf16 *x;     // x: [ncols][dim]
  int dim, vocab, codebook_size, num_codebooks, sem_begin, sem_end, scale_cb, cpu_sem, ncols;
  float inv_sqrt, sqrt_c;
  TokSrc tok; int *err;
};
template <class B> __device__ __forceinline__ void b_embed_body(const BEmbedArgs &a, int bx, int by, int bz, int gy, unsigned char *dsm) {
  (void)bx; (void)by; (void)bz; (void)gy; (void)dsm;
  const int n = bx;
  int tok = a.tok.base[n * a.tok.sn];
  if (tok < 0 || tok >= a.vocab) { tok = 0; if (threadIdx.x == 0) atomicExch(a.err, 1); }
  const bool is_sem = tok >= a.sem_begin && tok <= a.sem_end;
  for (int c = threadIdx.x; c * 8 < a.dim; c += B::nthreads()) {
    float te[8], vq[8];
    unpack8(*reinterpret_cast<const uint4 *>(a.emb + (size_t)tok * a.dim + c * 8), te);
#pragma unroll
    for (int j = 0; j < 8; ++j) vq[j] = 0.f;
    if (is_sem) {
      for (int i = 0; i < a.num_codebooks; ++i) {       // stack(...).sum(dim=1): fp32 accumulate, one rounding
        int cc = a.tok.base[n * a.tok.sn + (i + 1) * a.tok.sr];
        if (cc < 0 || cc >= a.codebook_size) { cc = 0; atomicExch(a.err, 1); }
        float ce[8]; unpack8(*reinterpret_cast<const uint4 *>(a.cb_emb + ((size_t)cc + (size_t)i * a.codebook_size) * a.dim + c * 8), ce);
#pragma unroll
        for (int j = 0; j < 8; ++j) vq[j] += ce[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) vq[j] = rbf(vq[j]);
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = rbf(te[j] + vq[j]);
      if (a.scale_cb && is_sem) x = a.cpu_sem ? rbf(__fdiv_rn(x, a.sqrt_c)) : rbf(__fmul_rn(x, a.inv_sqrt));
      o[j] = x;
    }
    uint4 u;
    u.x = (uint32_t)f2bits(o[0]) | ((uint32_t)f2bits(o[1]) << 16); u.y = (uint32_t)f2bits(o[2]) | ((uint32_t)f2bits(o[3]) << 16);
    u.z = (uint32_t)f2bits(o[4]) | ((uint32_t)f2bits(o[5]) << 16); u.w = (uint32_t)f2bits(o[6]) | ((uint32_t)f2bits(o[7]) << 16);
    *reinterpret_cast<uint4 *>(a.x + (size_t)n * a.dim + c * 8) = u;
  }
}
__global__ void __launch_bounds__(128) b_embed_kernel(const BEmbedArgs a) {
  extern __shared__ __align__(128) unsigned char dsm_b_embed_body[];
  pdl_launch_dependents();
  pdl_wait();
  b_embed_body<BlockAll>(a, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.y, dsm_b_embed_body);
}

// ---- RMSNorm: one warp per column -------------------------------------------------------------------------------------------------
struct BNormArgs { const bf16 *x, *w; bf16 *out; int K, ncols; float eps; };
template <class B> __device__ __forceinline__ void b_rmsnorm_body(const BNormArgs &a, int bx, int by, int bz, int gy, unsigned char *dsm) {
  (void)bx; (void)by; (void)bz; (void)gy; (void)dsm;
  const int lane = threadIdx.x & 31, n = bx * (B::nthreads() >> 5) + (threadIdx.x >> 5);
  if (n >= a.ncols) return;
  const uint4 *src = reinterpret_cast<const uint4 *>(a.x + (size_t)n * a.K);
  const int nc = a.K >> 3;
  float ss = 0.f;
  for (int c = lane; c < nc; c += 32) {
    float f[8]; unpack8(src[c], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss * (1.0f / (float)a.K) + a.eps);
  for (int c = lane; c < nc; c += 32) {
    float f[8], g[8]; unpack8(src[c], f); unpack8(reinterpret_cast<const uint4 *>(a.w)[c], g);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = rbf(__fmul_rn(rbf(__fmul_rn(f[j], inv)), g[j]));      // .type_as(x), then * weight
    uint4 u;
    u.x = (uint32_t)f2bits(f[0]) | ((uint32_t)f2bits(f[1]) << 16); u.y = (uint32_t)f2bits(f[2]) | ((uint32_t)f2bits(f[3]) << 16);
    u.z = (uint32_t)f2bits(f[4]) | ((uint32_t)f2bits(f[5]) << 16); u.w = (uint32_t)f2bits(f[6]) | ((uint32_t)f2bits(f[7]) << 16);
    reinterpret_cast<uint4 *>(a.out + (size_t)n * a.K)[c] = u;
  }
}
__global__ void __launch_bounds__(128) b_rmsnorm_kernel(const BNormArgs a) {
  extern __shared__ __align__(128) unsigned char dsm_b_rmsnorm_body[];
  pdl_launch_dependents();
  pdl_wait();
  b_rmsnorm_body<BlockAll>(a, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.y, dsm_b_rmsnorm_body);
}

// ---- q/k norm + RoPE in place, K/V rows into the cache ----------------------------------------------------------------------------
struct BQkvPostArgs {
  bf16 *qkv;                 // [ncols][(nh + 2 nkv) * hd]; q is rewritten in place
  bf16 *kc, *vc;             // cache of column n at + n * slot_stride elements: [nkv][S][hd]
  long long slot_stride;
  const bf16 *rope, *qn, *kn;
  int nh, nkv, hd, S, ncols; float eps;
  PosSrc pos;
};
template <class B> __device__ __forceinline__ void b_qkv_post_body(const BQkvPostArgs &a, int bx, int by, int bz, int gy, unsigned char *dsm) {
  (void)bx; (void)by; (void)bz; (void)gy; (void)dsm;
  float *sm_qp = reinterpret_cast<float *>(dsm);      // one head vector (hd floats) per warp
  const int n = bx, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = B::nthreads() >> 5;
  const int hd = a.hd, qd = a.nh * hd, kd = a.nkv * hd;
  const int pos = pos_of(a.pos, n);
  if (pos < 0 || pos >= a.S) return;
  bf16 *row = a.qkv + (size_t)n * (qd + 2 * kd);
  bf16 *kc = a.kc + (size_t)n * a.slot_stride, *vc = a.vc + (size_t)n * a.slot_stride;
  const bf16 *rope_row = a.rope + (size_t)pos * hd;
  float *v = sm_qp + (size_t)w * hd;
  // grid (ncols, head groups): one head vector per warp, so the dependent round trips of a head run in parallel across heads
  for (int h = by * nw + w; h < a.nh + 2 * a.nkv; h += nw * gy) {
    bf16 *src = row + (size_t)h * hd;
    if (h >= a.nh + a.nkv) {            // v head: KVCache.update (llama.py:142-149), no norm / rotation
      const int g = h - a.nh - a.nkv;
      for (int d = lane; d < hd; d += 32) vc[((size_t)g * a.S + pos) * hd + d] = src[d];
      continue;
    }
    for (int d = lane; d < hd; d += 32) v[d] = bf2f(src[d]);
    __syncwarp();
    head_norm_rope(v, hd, h < a.nh ? a.qn : a.kn, a.eps, rope_row, lane);
    if (h < a.nh) { for (int d = lane; d < hd; d += 32) src[d] = f2bf(v[d]); }
    else { const int g = h - a.nh; for (int d = lane; d < hd; d += 32) kc[((size_t)g * a.S + pos) * hd + d] = f2bf(v[d]); }
    __syncwarp();
  }
}
__global__ void __launch_bounds__(256) b_qkv_post_kernel(const BQkvPostArgs a) {
  extern __shared__ __align__(128) unsigned char dsm_b_qkv_post_body[];
  pdl_launch_dependents();
  pdl_wait();
  b_qkv_post_body<BlockAll>(a, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.y, dsm_b_qkv_post_body);
}

// ---- slow attention: split-KV flash-decode, one query per column ----------------------------------------------------------------
struct BAttnArgs {
  const bf16 *qkv;           // [ncols][(nh + 2 nkv) * hd]; fuse_post = 0: q already normalised + rotated (b_qkv_post_kernel ran)
  bf16 *kc, *vc; long long slot_stride;
  // fuse_post = 1 (batched decode; one cache per column): the raw wqkv output comes in and the CTA applies q/k nn.RMSNorm + RoPE itself
  // (llama.py:246-251); the CTA whose tiles contain the new position writes the K / V row into the cache (KVCache.update, :142-149)
  int fuse_post; const bf16 *rope, *qn, *kn; float eps;
  // cluster_merge = 1: the KV splits of a (column, kv head) are a thread-block cluster (cluster dims (nsplit_max, 1, 1)); their partial
  // (max, sum, output) meet in the first split's shared memory instead of a global buffer + ticket
  int cluster_merge;
  int use_mma;               // Q.K^T and P@V on the tensor cores (b_attn_body<B, true>)
  int nh, nkv, hd, S, ncols, nsplit_max, tiles_per_split; float sf;
  float *part_o, *part_ml;   // [ncols][nkv][nsplit_max][G][hd], [...][G][2]
  unsigned int *tickets;     // [ncols][nkv], zero between launches
  bf16 *y;                   // [ncols][nh * hd]
  PosSrc pos; int *err;
};
// grid (nsplit_max, nkv, ncols); dynamic smem b_attn_smem().  CTA (kv head, split) walks its 64-position tiles (bulk copies, double
// buffered).  Inside the CTA the positions of a tile are dealt to the 8 warps (rows w, w + 8, ...): every warp keeps its own running
// (max, sum, output) for two query heads at a time -- scores with 8 lanes per position (a 3-step reduction and one exp per lane),
// probabilities handed to all lanes for P@V where a lane owns hd/32 output dims -- and the warps meet ONCE, after the last tile
// (the decode kernel's tile walk, mega.cuh).  One CTA barrier per tile instead of three, no idle warps during the softmax.
#define DA_B_AWARPS (DA_ATTN_THREADS / 32)
#define DA_B_MAXSPLIT 4   // KV splits per (column, kv head) at most (a cluster; the receive buffer must leave room for two CTAs per SM)
#define DA_B_NBUF 3      // K/V tile buffers: two tiles are in flight while one is being consumed (a tile's HBM latency is ~2x its compute time)
static_assert(DA_TILE == 8 * DA_B_AWARPS, "a warp owns 8 positions of a tile: 4 lanes each");
static inline size_t b_attn_smem(int G, int hd) {
  size_t f = ((size_t)G * hd + (size_t)DA_B_AWARPS * G * (2 + hd)) * sizeof(float) + 2 * sizeof(uint64_t) + (size_t)hd * (sizeof(float) + 2 * sizeof(bf16)) +
             (size_t)DA_B_MAXSPLIT * G * (hd + 2) * sizeof(float);
  f = (f + 127) & ~(size_t)127;
  return f + 128 + 2 * DA_B_NBUF * (size_t)DA_TILE * hd * sizeof(bf16);
}
// `bars`: DA_B_NBUF mbarriers in STATIC shared memory; `phase`: their parities as this thread has seen them.  The stand-alone kernel
// initialises them per launch (init = true); the persistent kernel initialises them once and carries the parities from unit to unit --
// re-initialising an mbarrier in the middle of a kernel (through a per-thread address: a plain 64-bit store + a sync-unit cache
// invalidate in SASS) lost the transaction that followed on B200.
// MMA = true: Q.K^T and P@V on the tensor cores (mma.sync m16n8k16, bf16 in, fp32 accumulate).  The G <= 8 query heads of the kv head are
// rows 0 .. G-1 of the 16-row A operand.  Q and K are bf16 values, so every product q_d * k_d is exact in fp32 and the dot product times
// scale (= sqrt(scale)^2) differs from the reference's sum of (q_d sqrt(scale)) (k_d sqrt(scale)) by fp32 roundings only.  P is fp32: it
// is split EXACTLY into three bf16 terms (24 significant bits = 3 x 8) that accumulate into the same fp32 output registers, so P@V is the
// fp32 product as well.  K / V tiles come in by 2-D TMA (cp.async.bulk.tensor) as boxes of 64 positions x 64 dims with the 128-byte
// swizzle, so the 32-bit fragment loads of K and the transposing ldmatrix of V are bank-conflict free (one bulk copy PER ROW at a padded
// stride was tried first: 128 small copies per tile cost more than the arithmetic they saved).  A warp owns 16 positions of a 64-position
// tile (w % 4) and half of the output dims (w / 4); the two warps of a position block compute the same scores.  hd = 64 or 128.
#define DA_BM_NBUF 3
static inline size_t b_attn_mma_smem(int G, int hd) {
  size_t f = ((size_t)G * hd + (size_t)DA_B_AWARPS * G * (2 + hd)) * sizeof(float) + 2 * sizeof(uint64_t) + (size_t)hd * (sizeof(float) + 2 * sizeof(bf16)) +
             (size_t)DA_B_MAXSPLIT * G * (hd + 2) * sizeof(float);
  f = (f + 127) & ~(size_t)127;
  return f + 1024 + 2 * DA_BM_NBUF * (size_t)DA_TILE * hd * sizeof(bf16);
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { return (uint32_t)f2bits(lo) | ((uint32_t)f2bits(hi) << 16); }
// element (row r, dim d) of a tile made of 64-dim boxes [64 rows][128 bytes] with the 128-byte swizzle: byte offset
__device__ __forceinline__ uint32_t sw128_off(int r, int d) {
  return (uint32_t)((d >> 6) * (DA_TILE * 128) + r * 128 + ((((d & 63) >> 3) ^ (r & 7)) << 4) + (d & 7) * 2);
}
template <class B, bool MMA = false, int HD = 128> __device__ __forceinline__ void b_attn_body(const BAttnArgs &a, int bx, int by, int bz, int gy, unsigned char *dsm,
                                                               uint64_t *bars, uint32_t (&phase)[DA_B_NBUF], bool init,
                                                               const CUtensorMap *mk = nullptr, const CUtensorMap *mv = nullptr) {
  constexpr int NBUF = MMA ? DA_BM_NBUF : DA_B_NBUF;
  (void)bx; (void)by; (void)bz; (void)gy; (void)dsm;
  unsigned char *smraw_b = dsm;
  const int g = by, split = bx, n = bz;
  const int G = a.nh / a.nkv, hd = a.hd, dpl = hd >> 5;      // output dims per lane (1, 2 or 4)
  const int pos = pos_of(a.pos, n);
  if (pos < 0 || pos >= a.S) return;
  const int L = pos + 1;
  const int n_tiles = (L + DA_TILE - 1) / DA_TILE;
  // a split is worth its partials / ticket / merge only from tiles_per_split tiles on (8 = 512 positions by default): measured at
  // 4 x 32 slots, context ~360, 6.73 ms per round with 8 tiles per split against 7.04 with 4
  const int nsplit = max(1, min(a.nsplit_max, (n_tiles + a.tiles_per_split - 1) / a.tiles_per_split));
  const int tps = (n_tiles + nsplit - 1) / nsplit;
  const int nsplit_eff = (n_tiles + tps - 1) / tps;
  const bool cl = a.cluster_merge != 0;
  // cluster barrier 1 (arrive here, wait before the first remote store) proves that the first split's CTA is running; barrier 2 publishes
  // the partials.  A CTA without tiles takes part in both and leaves; a column without a request is skipped by its whole cluster alike.
  if (cl) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  if (split >= nsplit_eff) {
    if (cl) { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
    return;
  }
  const int t0 = split * tps, t1 = min(n_tiles, t0 + tps);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *q = reinterpret_cast<float *>(smraw_b);
  float *pm = q + G * hd, *pl = pm + DA_B_AWARPS * G, *po = pl + DA_B_AWARPS * G;      // per-warp partials: [8][G], [8][G], [8][G][hd]
  float *knew_f = po + (size_t)DA_B_AWARPS * G * hd;      // fuse_post: the new K row while it is normalised (fp32), then the new K and V rows (bf16)
  bf16 *krow_new = reinterpret_cast<bf16 *>(knew_f + hd), *vrow_new = krow_new + hd;
  float *recv = reinterpret_cast<float *>(vrow_new + hd);      // cluster merge: [split][G * hd outputs | G x (max, sum)]
  const int rstride = G * (hd + 2);
  size_t off = (size_t)((unsigned char *)(recv + (size_t)DA_B_MAXSPLIT * rstride) - smraw_b) + 16;
  if (MMA) off += (1024u - ((smem_u32(smraw_b) + (uint32_t)off) & 1023u)) & 1023u;      // swizzled boxes are 1024-byte aligned
  else off = (off + 127) & ~(size_t)127;
  bf16 *kbuf = reinterpret_cast<bf16 *>(smraw_b + off);
  const int rs = hd;
  bf16 *vbuf = kbuf + (size_t)NBUF * DA_TILE * rs;
  const bf16 *kc = a.kc + (size_t)n * a.slot_stride, *vc = a.vc + (size_t)n * a.slot_stride;
  const uint64_t pol = policy_evict_first();
  if (init && threadIdx.x == 0) {
    for (int i = 0; i < NBUF; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  B::sync();
  auto issue = [&](int t, int buf) {      // scalar path: thread 0; MMA path: every lane of warp 0 (one bulk copy per row, padded stride)
    const int r0 = t * DA_TILE, r1 = min(L, r0 + DA_TILE);
    const uint32_t bytes = (uint32_t)(r1 - r0) * hd * sizeof(bf16);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (!MMA) {
      mbar_expect_tx(&bars[buf], 2 * bytes);
      bulk_g2s(kbuf + (size_t)buf * DA_TILE * hd, kc + ((size_t)g * a.S + r0) * hd, bytes, &bars[buf], pol);
      bulk_g2s(vbuf + (size_t)buf * DA_TILE * hd, vc + ((size_t)g * a.S + r0) * hd, bytes, &bars[buf], pol);
    } else {
      // whole 64-row boxes (rows past the context are other positions' finite values or zero fill: masked below), hd / 64 boxes per tile
      mbar_expect_tx(&bars[buf], 2u * DA_TILE * (uint32_t)hd * sizeof(bf16));
      const int grow = (int)(((long long)n * a.slot_stride) / hd) + g * a.S + r0;
      for (int bxi = 0; bxi < (hd >> 6); ++bxi) {
        tma_load_2d(reinterpret_cast<unsigned char *>(kbuf + (size_t)buf * DA_TILE * hd) + bxi * (DA_TILE * 128), mk, bxi * 64, grow, &bars[buf], pol);
        tma_load_2d(reinterpret_cast<unsigned char *>(vbuf + (size_t)buf * DA_TILE * hd) + bxi * (DA_TILE * 128), mv, bxi * 64, grow, &bars[buf], pol);
      }
    }
  };
  if (threadIdx.x == 0) { issue(t0, 0); if (t0 + 1 < t1) issue(t0 + 1, 1); }
  const int t_new = pos / DA_TILE;
  const bool owns_new = a.fuse_post && t_new >= t0 && t_new < t1;      // this CTA's tiles contain the position being decoded
  {
    const bf16 *row = a.qkv + (size_t)n * (a.nh + 2 * a.nkv) * hd;
    const bf16 *qsrc = row + (size_t)g * G * hd;
    for (int c = threadIdx.x; c * 8 < G * hd; c += DA_ATTN_THREADS) {
      float t[8]; unpack8(*reinterpret_cast<const uint4 *>(qsrc + c * 8), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) q[c * 8 + j] = (a.fuse_post || MMA) ? t[j] : __fmul_rn(t[j], a.sf);      // q * sqrt(scale), fp32 (math SDPA)
    }
    if (a.fuse_post) {
      if (owns_new) for (int d = threadIdx.x; d < hd; d += DA_ATTN_THREADS) knew_f[d] = bf2f(row[(size_t)(a.nh + g) * hd + d]);
      B::sync();
      // tasks 0 .. G-1: the query heads of this kv head; G: the new K row; G + 1: the new V row -- one warp each
      const bf16 *rope_row = a.rope + (size_t)pos * hd;
      for (int task = w; task < G + (owns_new ? 2 : 0); task += DA_B_AWARPS) {
        if (task < G) {
          head_norm_rope(q + (size_t)task * hd, hd, a.qn, a.eps, rope_row, lane);
          if (!MMA) for (int d = lane; d < hd; d += 32) q[(size_t)task * hd + d] = __fmul_rn(q[(size_t)task * hd + d], a.sf);
        } else if (task == G) {
          head_norm_rope(knew_f, hd, a.kn, a.eps, rope_row, lane);
          for (int d = lane; d < hd; d += 32) { const bf16 kv = f2bf(knew_f[d]); krow_new[d] = kv; a.kc[(size_t)n * a.slot_stride + ((size_t)g * a.S + pos) * hd + d] = kv; }
        } else {
          for (int d = lane; d < hd; d += 32) { const bf16 vv = row[(size_t)(a.nh + a.nkv + g) * hd + d]; vrow_new[d] = vv; a.vc[(size_t)n * a.slot_stride + ((size_t)g * a.S + pos) * hd + d] = vv; }
        }
      }
    }
  }
  B::sync();
  bool ok = true;
  if constexpr (!MMA) {
  // Per tile a warp owns the 8 positions w, w + 8, ...: scores with 4 lanes per position (lane = 4 * slot + dl; every lane walks its
  // hd / 4 dims in 4-element chunks, the chunk order rotated by the slot so that the four positions of a half-warp hit different banks),
  // ONE online-softmax update per tile and warp, then P@V with the 8 probabilities broadcast to all lanes (a lane owns hd / 32 output
  // dims).  The query is pre-multiplied by sqrt(scale) and the dot product by sqrt(scale) again -- the reference scales q and k
  // separately (llama.py:265-270 through the math SDPA backend); the difference is one fp32 rounding.  With G <= 2 (one head pair)
  // the running (max, sum, output) stay in registers for the CTA's whole walk.
  const int psl = lane >> 2, dl = lane & 3, nit = hd >> 4;
  const bool single = G <= 2;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f}, o_acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int t = t0; t < t1; ++t) {
    const int buf = (t - t0) % DA_B_NBUF;
    const int r0 = t * DA_TILE, r1 = min(L, r0 + DA_TILE), nrow = r1 - r0;
    if (threadIdx.x == 0 && t + 2 < t1) issue(t + 2, (t - t0 + 2) % DA_B_NBUF);      // into the buffer of tile t - 1, released by the barrier that ended it
    ok = mbar_wait(&bars[buf], phase[buf]) && ok; phase[buf] ^= 1u;
    bf16 *kt = kbuf + (size_t)buf * DA_TILE * hd, *vt = vbuf + (size_t)buf * DA_TILE * hd;
    if (owns_new && t == t_new) {
      // the bulk copy brought whatever the cache held at row `pos`; the warp that owns that row of the tile (rows w, w + 8, ...) replaces
      // it with the row computed above -- only this warp reads it, in Q.K and in P@V, so a warp-level fence is all it takes
      const int jrow = pos - r0;
      if ((jrow & (DA_B_AWARPS - 1)) == w) {
        for (int d = lane; d < hd; d += 32) { kt[(size_t)jrow * hd + d] = krow_new[d]; vt[(size_t)jrow * hd + d] = vrow_new[d]; }
        __syncwarp();
      }
    }
    for (int h0 = 0; h0 < G; h0 += 2) {
      const bool first = t == t0;
      if (!single) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const bool hv = h0 + hh < G;
          const int h = hv ? h0 + hh : h0;
          m_run[hh] = first ? -INFINITY : pm[w * G + h]; l_run[hh] = first ? 0.f : pl[w * G + h];
#pragma unroll
          for (int i = 0; i < 4; ++i) o_acc[hh][i] = (hv && i < dpl && !first) ? po[((size_t)w * G + h) * hd + lane * dpl + i] : 0.f;
        }
      }
      const float *q0 = q + (size_t)h0 * hd, *q1 = q + (size_t)(h0 + 1 < G ? h0 + 1 : h0) * hd;
      const int jme = w + DA_B_AWARPS * psl;
      float s2[2] = {0.f, 0.f};
      if (jme < nrow) {
        const bf16 *krow = kt + (size_t)jme * hd;
#pragma unroll 4
        for (int i = 0; i < nit; ++i) {
          const int e = dl * 4 + 16 * ((i + psl) & (nit - 1));
          float kf[4];
          { const uint2 u = *reinterpret_cast<const uint2 *>(krow + e);
            kf[0] = __uint_as_float(u.x << 16); kf[1] = __uint_as_float(u.x & 0xffff0000u); kf[2] = __uint_as_float(u.y << 16); kf[3] = __uint_as_float(u.y & 0xffff0000u); }
          const float4 qa = *reinterpret_cast<const float4 *>(q0 + e), qb = *reinterpret_cast<const float4 *>(q1 + e);
          s2[0] = fmaf(qa.x, kf[0], s2[0]); s2[0] = fmaf(qa.y, kf[1], s2[0]); s2[0] = fmaf(qa.z, kf[2], s2[0]); s2[0] = fmaf(qa.w, kf[3], s2[0]);
          s2[1] = fmaf(qb.x, kf[0], s2[1]); s2[1] = fmaf(qb.y, kf[1], s2[1]); s2[1] = fmaf(qb.z, kf[2], s2[1]); s2[1] = fmaf(qb.w, kf[3], s2[1]);
        }
      }
      float pj[DA_B_AWARPS][2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float sv = s2[hh];
        sv += __shfl_xor_sync(0xffffffffu, sv, 1); sv += __shfl_xor_sync(0xffffffffu, sv, 2);
        sv = (jme < nrow && h0 + hh < G) ? __fmul_rn(sv, a.sf) : -INFINITY;
        float mx = fmaxf(sv, __shfl_xor_sync(0xffffffffu, sv, 4));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        const float m_new = fmaxf(m_run[hh], mx);
        float pme = 0.f;
        if (m_new != -INFINITY) {
          const float sc_old = expf(m_run[hh] - m_new);      // exp(-inf) = 0 before the first position
          l_run[hh] *= sc_old; m_run[hh] = m_new;
#pragma unroll
          for (int i = 0; i < 4; ++i) o_acc[hh][i] *= sc_old;
          pme = expf(sv - m_new);
        }
#pragma unroll
        for (int jj = 0; jj < DA_B_AWARPS; ++jj) pj[jj][hh] = __shfl_sync(0xffffffffu, pme, jj * 4);
      }
#pragma unroll
      for (int jj = 0; jj < DA_B_AWARPS; ++jj) {
        const int j = w + DA_B_AWARPS * jj;
        if (j < nrow) {
          float vf[4] = {0.f, 0.f, 0.f, 0.f};
          const bf16 *vp = vt + (size_t)j * hd + lane * dpl;
          if (dpl == 4) { const uint2 u = *reinterpret_cast<const uint2 *>(vp); vf[0] = __uint_as_float(u.x << 16); vf[1] = __uint_as_float(u.x & 0xffff0000u); vf[2] = __uint_as_float(u.y << 16); vf[3] = __uint_as_float(u.y & 0xffff0000u); }
          else if (dpl == 2) { const uint32_t u = *reinterpret_cast<const uint32_t *>(vp); vf[0] = __uint_as_float(u << 16); vf[1] = __uint_as_float(u & 0xffff0000u); }
          else vf[0] = bf2f(*vp);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            l_run[hh] += pj[jj][hh];
#pragma unroll
            for (int i = 0; i < 4; ++i) o_acc[hh][i] = fmaf(pj[jj][hh], vf[i], o_acc[hh][i]);
          }
        }
      }
      if (!single || t + 1 == t1) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (h0 + hh < G) {
            const int h = h0 + hh;
            if (lane == 0) { pm[w * G + h] = m_run[hh]; pl[w * G + h] = l_run[hh]; }
#pragma unroll
            for (int i = 0; i < 4; ++i) if (i < dpl) po[((size_t)w * G + h) * hd + lane * dpl + i] = o_acc[hh][i];
          }
        }
        __syncwarp();
      }
    }
    B::sync();      // every warp is past its last read of the tile: its buffer may be refilled
  }
  } else {
    // ---- tensor-core tile walk ----------------------------------------------------------------------------------------------------
    constexpr int hh = HD / 2, nkk = HD / 16, nnt = HD / 16;      // dims per half, k-steps, n-tiles per half (HD = head_dim, 64 or 128)
    const int pb = w & 3, dh = w >> 2;                           // position block, dim half
    const int hrow = lane >> 2, qc = (lane & 3) * 2;                                       // this lane's head row / column pair inside a fragment
    const float scale = __fmul_rn(a.sf, a.sf);
    uint32_t qa0[nkk], qa2[nkk];      // A fragments of Q: rows >= G are zero
#pragma unroll
    for (int kk = 0; kk < nkk; ++kk) {
      qa0[kk] = 0u; qa2[kk] = 0u;
      if (hrow < G) {
        const float *qr = q + (size_t)hrow * hd + kk * 16 + qc;
        qa0[kk] = pack_bf16x2(qr[0], qr[1]); qa2[kk] = pack_bf16x2(qr[8], qr[9]);
      }
    }
    float m_run = -INFINITY, l_run = 0.f, o[nnt][4];
#pragma unroll
    for (int j = 0; j < nnt; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    // swizzled fragment addresses: element (row r, dim d) lives at box (d / 64), row r, 16-byte chunk ((d % 64) / 8) ^ (r % 8).  The rows a
    // lane touches keep r % 8 for the whole walk (K: lane / 4, V: lane % 8), so the XOR term is a per-lane constant and a chunk index that
    // is known at compile time costs one LOP3
    const uint32_t kx = (uint32_t)(hrow & 7) << 4, vx = (uint32_t)(lane & 7) << 4;
    for (int t = t0; t < t1; ++t) {
      const int buf = (t - t0) % NBUF;
      const int r0 = t * DA_TILE, r1 = min(L, r0 + DA_TILE), nrow = r1 - r0;
      if (threadIdx.x == 0 && t + 2 < t1) issue(t + 2, (t - t0 + 2) % NBUF);      // into the buffer of tile t - 1, released by the barrier that ended it
      ok = mbar_wait(&bars[buf], phase[buf]) && ok; phase[buf] ^= 1u;
      unsigned char *kt = reinterpret_cast<unsigned char *>(kbuf + (size_t)buf * DA_TILE * hd), *vt = reinterpret_cast<unsigned char *>(vbuf + (size_t)buf * DA_TILE * hd);
      const int p0 = pb * 16;      // this warp's 16 positions of the tile
      // the new position's row is not in the cache yet when the tile is fetched: both warps of its position block patch it in (same values)
      if (owns_new && t == t_new) {
        const int jrow = pos - r0;
        if (jrow >= p0 && jrow < p0 + 16)
          for (int d = lane; d < hd; d += 32) {
            *reinterpret_cast<bf16 *>(kt + sw128_off(jrow, d)) = krow_new[d]; *reinterpret_cast<bf16 *>(vt + sw128_off(jrow, d)) = vrow_new[d];
          }
      }
      __syncwarp();
      // S = Q K^T for the two 8-position n-tiles of the block
      float sacc[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f;
        const unsigned char *kr = kt + (p0 + j * 8 + hrow) * 128 + qc * 2;      // B fragment: position = lane / 4, dims (lane % 4) * 2 + {0, 1} (+ 8)
#pragma unroll
        for (int kk = 0; kk < nkk; ++kk) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t *>(kr + (kk >> 2) * (DA_TILE * 128) + ((uint32_t)(((2 * kk) & 7) << 4) ^ kx));
          const uint32_t b1 = *reinterpret_cast<const uint32_t *>(kr + (kk >> 2) * (DA_TILE * 128) + ((uint32_t)(((2 * kk + 1) & 7) << 4) ^ kx));
          mma_bf16_16816(sacc[j], qa0[kk], 0u, qa2[kk], 0u, b0, b1);
        }
      }
      // online softmax for head row `hrow` over this warp's 16 positions: the lane holds 4 of them, its 3 neighbours the rest
      float sv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pcol = p0 + (i >> 1) * 8 + qc + (i & 1);
        sv[i] = (pcol < nrow && hrow < G) ? __fmul_rn(sacc[i >> 1][i & 1], scale) : -INFINITY;
      }
      float mx = fmaxf(fmaxf(sv[0], sv[1]), fmaxf(sv[2], sv[3]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);
      float pr[4] = {0.f, 0.f, 0.f, 0.f}, sc_old = 1.f;
      if (m_new != -INFINITY) {
        sc_old = expf(m_run - m_new);      // exp(-inf) = 0 before the first position
        m_run = m_new;
#pragma unroll
        for (int i = 0; i < 4; ++i) pr[i] = expf(sv[i] - m_new);
      }
      float rsum = (pr[0] + pr[1]) + (pr[2] + pr[3]);
      rsum += __shfl_xor_sync(0xffffffffu, rsum, 1); rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
      l_run = fmaf(l_run, sc_old, rsum);
      // P as three exact bf16 terms; A fragments: a0 = columns of n-tile 0 (k 0..7), a2 = n-tile 1 (k 8..15)
      uint32_t pa0[3], pa2[3];
      {
        float r0f[4], hi;
#pragma unroll
        for (int i = 0; i < 4; ++i) r0f[i] = pr[i];
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) {
          float tt[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { hi = rbf(r0f[i]); tt[i] = hi; r0f[i] = r0f[i] - hi; }
          pa0[s3] = pack_bf16x2(tt[0], tt[1]); pa2[s3] = pack_bf16x2(tt[2], tt[3]);
        }
      }
      // O = O * sc_old + P V over this warp's dim half: V^T fragments by transposing ldmatrix (two n-tiles per instruction)
      // this lane's row address for ldmatrix: row p0 + (lane / 8 % 2) * 8 + lane % 8, dims dh * hh + (lane / 16) * 8 + j * 8
      const int vd0 = dh * hh + (lane >> 4) * 8;
      const uint32_t vr = smem_u32(vt) + (uint32_t)((p0 + ((lane >> 3) & 1) * 8 + (lane & 7)) * 128 + (vd0 >> 6) * (DA_TILE * 128));
      const uint32_t vc0 = (uint32_t)((vd0 & 63) >> 3);      // chunk of n-tile 0 inside its box; n-tile j adds j (hh <= 64: never leaves the box)
#pragma unroll
      for (int j = 0; j < nnt; j += 2) {
        uint32_t b00, b01, b10, b11;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(b00), "=r"(b01), "=r"(b10), "=r"(b11) : "r"(vr + ((((vc0 + (uint32_t)j) & 7u) << 4) ^ vx)));
        o[j][0] *= sc_old; o[j][1] *= sc_old; o[j + 1][0] *= sc_old; o[j + 1][1] *= sc_old;      // rows 8 .. 15 of the fragments stay zero
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) {
          mma_bf16_16816(o[j], pa0[s3], 0u, pa2[s3], 0u, b00, b01);
          mma_bf16_16816(o[j + 1], pa0[s3], 0u, pa2[s3], 0u, b10, b11);
        }
      }
      B::sync();      // every warp is past its last read of the tile: its buffer may be refilled
    }
    // per-warp partials for the fold: (max, sum) per head, output for this warp's dim half
    if (hrow < G) {
      if ((lane & 3) == 0) { pm[w * G + hrow] = m_run; pl[w * G + hrow] = l_run; }
#pragma unroll
      for (int j = 0; j < nnt; ++j) { po[((size_t)w * G + hrow) * hd + dh * hh + j * 8 + qc] = o[j][0]; po[((size_t)w * G + hrow) * hd + dh * hh + j * 8 + qc + 1] = o[j][1]; }
    }
    B::sync();
  }
  if (!ok && threadIdx.x == 0) atomicExch(a.err, 2);
  if (cl) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");      // barrier 1: every CTA of the cluster is running
  // fold the 8 warps' partials in warp order: thread e = (h, d)
  const size_t pbase = ((size_t)n * a.nkv + g) * a.nsplit_max;
  float *pog = a.part_o + ((pbase + split) * G) * hd;
  float *pml = a.part_ml + ((pbase + split) * G) * 2;
  for (int e = threadIdx.x; e < G * hd; e += DA_ATTN_THREADS) {
    const int h = e / hd, dd = e - h * hd;
    // scalar path: all 8 warps hold partials of every dim; tensor-core path: the 4 warps (position blocks) of this dim's half
    const int w_lo = MMA ? (dd >= (hd >> 1) ? 4 : 0) : 0, w_hi = MMA ? w_lo + 4 : DA_B_AWARPS;
    float m = -INFINITY;
    for (int ww = w_lo; ww < w_hi; ++ww) m = fmaxf(m, pm[ww * G + h]);
    float l = 0.f, o = 0.f;
    for (int ww = w_lo; ww < w_hi; ++ww) {
      const float v = pm[ww * G + h];
      const float sc_w = v == -INFINITY ? 0.f : expf(v - m);
      l = fmaf(pl[ww * G + h], sc_w, l);
      o = fmaf(po[((size_t)ww * G + h) * hd + dd], sc_w, o);
    }
    if (nsplit_eff == 1) a.y[(size_t)n * a.nh * hd + (size_t)g * G * hd + e] = f2bf(o / l);      // single split: no partials, no ticket
    else if (!cl) { pog[e] = o; if (dd == 0) { pml[h * 2] = m; pml[h * 2 + 1] = l; } }
    else {
      // into split 0's receive buffer (its own for split 0)
      uint32_t dst;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(smem_u32(recv + (size_t)split * rstride)), "r"(0));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst + (uint32_t)e * 4u), "f"(o) : "memory");
      if (dd == 0) {
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst + (uint32_t)(G * hd + h * 2) * 4u), "f"(m) : "memory");
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst + (uint32_t)(G * hd + h * 2 + 1) * 4u), "f"(l) : "memory");
      }
    }
  }
  if (cl) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    if (nsplit_eff == 1 || split != 0) return;
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    for (int e = threadIdx.x; e < G * hd; e += DA_ATTN_THREADS) {      // merge in split order: the arithmetic of the ticket path below
      const int h = e / hd;
      float m = -INFINITY;
      for (int s = 0; s < nsplit_eff; ++s) m = fmaxf(m, recv[(size_t)s * rstride + G * hd + h * 2]);
      float l = 0.f, o = 0.f;
      for (int s = 0; s < nsplit_eff; ++s) {
        const float *rs = recv + (size_t)s * rstride;
        const float sc_s = expf(rs[G * hd + h * 2] - m);
        l = fmaf(rs[G * hd + h * 2 + 1], sc_s, l);
        o = fmaf(rs[e], sc_s, o);
      }
      a.y[(size_t)n * a.nh * hd + (size_t)g * G * hd + e] = f2bf(o / l);
    }
    return;
  }
  if (nsplit_eff == 1) return;
  __shared__ unsigned int s_last;
  __threadfence();
  B::sync();
  unsigned int *ticket = a.tickets + (size_t)n * a.nkv + g;
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned)nsplit_eff - 1);
  B::sync();
  if (!s_last) return;
  __threadfence();
  for (int e = threadIdx.x; e < G * hd; e += DA_ATTN_THREADS) {      // merge in split order (deterministic)
    const int h = e / hd;
    float m = -INFINITY;
    for (int s = 0; s < nsplit_eff; ++s) m = fmaxf(m, __ldcg(a.part_ml + ((pbase + s) * G + h) * 2));
    float l = 0.f, o = 0.f;
    for (int s = 0; s < nsplit_eff; ++s) {
      const float *ml = a.part_ml + ((pbase + s) * G + h) * 2;
      const float sc_s = expf(__ldcg(ml) - m);
      l = fmaf(__ldcg(ml + 1), sc_s, l);
      o = fmaf(__ldcg(a.part_o + ((pbase + s) * G) * hd + e), sc_s, o);
    }
    a.y[(size_t)n * a.nh * hd + (size_t)g * G * hd + e] = f2bf(o / l);
  }
  if (threadIdx.x == 0) *ticket = 0u;
}
__global__ void __launch_bounds__(DA_ATTN_THREADS, 2) b_attn_kernel(const BAttnArgs a) {
  extern __shared__ __align__(128) unsigned char dsm_b_attn_body[];
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(8) uint64_t s_bars[DA_B_NBUF];
  uint32_t phase[DA_B_NBUF] = {};
  b_attn_body<BlockAll>(a, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.y, dsm_b_attn_body, s_bars, phase, true);
}
__global__ void __launch_bounds__(DA_ATTN_THREADS, 2) b_attn_mma_kernel(const __grid_constant__ CUtensorMap mk, const __grid_constant__ CUtensorMap mv, const BAttnArgs a) {
  extern __shared__ __align__(128) unsigned char dsm_b_attn_mma[];
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(8) uint64_t s_bars[DA_B_NBUF];
  uint32_t phase[DA_B_NBUF] = {};
  if (a.hd == 128) b_attn_body<BlockAll, true, 128>(a, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.y, dsm_b_attn_mma, s_bars, phase, true, &mk, &mv);
  else b_attn_body<BlockAll, true, 64>(a, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.y, dsm_b_attn_mma, s_bars, phase, true, &mk, &mv);
}

// ---- fast-layer attention for codebook position p, one CTA per column ---------------------------------------------------------------
struct BFastAttnArgs {
  const bf16 *qkv;           // [ncols][(nh + 2 nkv) * hd]
  bf16 *kc, *vc;             // per column at + n * slot_stride: [nkv][ncb][hd]  (the reference's fast KVCache layout)
  long long slot_stride;
  const bf16 *rope, *qn, *kn;
  int nh, nkv, hd, ncb, p, ncols; float eps, scale;
  bf16 *y;                   // [ncols][nh * hd]
};
// dynamic smem: q[qd] | k[ncb][kd] | v[ncb][kd] | pr[nh][ncb]  floats
template <class B> __device__ __forceinline__ void b_fast_attn_body(const BFastAttnArgs &f, int bx, int by, int bz, int gy, unsigned char *dsm) {
  (void)bx; (void)by; (void)bz; (void)gy; (void)dsm;
  float *sm_fa = reinterpret_cast<float *>(dsm);
  const int n = bx, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = B::nthreads() >> 5;
  const int qd = f.nh * f.hd, kd = f.nkv * f.hd, P = f.p + 1, G = f.nh / f.nkv;
  float *q = sm_fa, *ka = q + qd, *va = ka + f.ncb * kd, *pr = va + f.ncb * kd;
  const bf16 *row = f.qkv + (size_t)n * (qd + 2 * kd);
  bf16 *kc = f.kc + (size_t)n * f.slot_stride, *vc = f.vc + (size_t)n * f.slot_stride;
  for (int c = threadIdx.x; c * 8 < qd + 2 * kd; c += B::nthreads()) {
    float t[8]; unpack8(*reinterpret_cast<const uint4 *>(row + c * 8), t);
    const int e = c * 8;
    float *dst = e < qd ? q + e : (e < qd + kd ? ka + f.p * kd + (e - qd) : va + f.p * kd + (e - qd - kd));
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = t[j];
  }
  const int hd8 = f.hd >> 3, kd8 = kd >> 3;
  for (int c = threadIdx.x; c < f.p * kd8; c += B::nthreads()) {      // earlier positions from the cache
    const int j = c / kd8, r = c - j * kd8, g = r / hd8, d8 = r - g * hd8;
    const size_t src = ((size_t)g * f.ncb + j) * f.hd + (size_t)d8 * 8;
    float t[8];
    unpack8(*reinterpret_cast<const uint4 *>(kc + src), t);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) ka[c * 8 + jj] = t[jj];
    unpack8(*reinterpret_cast<const uint4 *>(vc + src), t);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) va[c * 8 + jj] = t[jj];
  }
  B::sync();
  const bf16 *rope_row = f.rope + (size_t)f.p * f.hd;
  for (int h = w; h < f.nh + f.nkv; h += nw) {
    if (h < f.nh) head_norm_rope(q + h * f.hd, f.hd, f.qn, f.eps, rope_row, lane);
    else head_norm_rope(ka + f.p * kd + (h - f.nh) * f.hd, f.hd, f.kn, f.eps, rope_row, lane);
  }
  B::sync();
  for (int e = threadIdx.x; e < kd; e += B::nthreads()) {               // KVCache.update (llama.py:142-149)
    const int g = e / f.hd, d = e - g * f.hd;
    const size_t dst = ((size_t)g * f.ncb + f.p) * f.hd + d;
    kc[dst] = f2bf(ka[f.p * kd + e]); vc[dst] = f2bf(va[f.p * kd + e]);
  }
  for (int t = threadIdx.x; t < f.nh * P; t += B::nthreads()) {         // bf16(q @ k^T), then bf16(* scale)   (llama.py:304)
    const int h = t / P, j = t - h * P, g = h / G;
    const float4 *qq = reinterpret_cast<const float4 *>(q + h * f.hd), *kk = reinterpret_cast<const float4 *>(ka + j * kd + g * f.hd);
    float acc = 0.f;
    for (int d = 0; d < (f.hd >> 2); ++d) {
      const float4 x = qq[d], y = kk[d];
      acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
    }
    pr[h * f.ncb + j] = rbf(__fmul_rn(rbf(acc), f.scale));
  }
  B::sync();
  // softmax in fp32, rounded to bf16 (llama.py:305-306; masked columns are exp(-inf) = 0).  One thread per (h, j)
  // (nh * ncb <= 256, check_config); every thread of a row walks it in the same order, so the row sum is identical
  float pval = 0.f;
  const int tt = threadIdx.x;
  if (tt < f.nh * P) {
    const int h = tt / P, j = tt - h * P;
    float m = -INFINITY;
    for (int jj = 0; jj < P; ++jj) m = fmaxf(m, pr[h * f.ncb + jj]);
    float sum = 0.f;
    for (int jj = 0; jj < P; ++jj) sum += expf(pr[h * f.ncb + jj] - m);
    pval = rbf(expf(pr[h * f.ncb + j] - m) / sum);
  }
  B::sync();
  if (tt < f.nh * P) { const int h = tt / P, j = tt - h * P; pr[h * f.ncb + j] = pval; }
  B::sync();
  bf16 *yo = f.y + (size_t)n * qd;
  for (int c = threadIdx.x; c * 8 < qd; c += B::nthreads()) {           // y = bf16(p @ v)   (llama.py:309)
    const int e = c * 8, h = e / f.hd, d = e - h * f.hd, g = h / G;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int jj = 0; jj < P; ++jj) {
      const float pj = pr[h * f.ncb + jj];
      const float *vv = va + jj * kd + g * f.hd + d;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(pj, vv[j], acc[j]);
    }
    uint4 u;
    u.x = (uint32_t)f2bits(acc[0]) | ((uint32_t)f2bits(acc[1]) << 16); u.y = (uint32_t)f2bits(acc[2]) | ((uint32_t)f2bits(acc[3]) << 16);
    u.z = (uint32_t)f2bits(acc[4]) | ((uint32_t)f2bits(acc[5]) << 16); u.w = (uint32_t)f2bits(acc[6]) | ((uint32_t)f2bits(acc[7]) << 16);
    *reinterpret_cast<uint4 *>(yo + e) = u;
  }
}
__global__ void __launch_bounds__(256) b_fast_attn_kernel(const BFastAttnArgs f) {
  extern __shared__ __align__(128) unsigned char dsm_b_fast_attn_body[];
  pdl_launch_dependents();
  pdl_wait();
  b_fast_attn_body<BlockAll>(f, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.y, dsm_b_fast_attn_body);
}
static inline size_t b_fast_attn_smem(int nh, int nkv, int hd, int ncb) {
  return ((size_t)nh * hd + 2 * (size_t)ncb * nkv * hd + (size_t)nh * ncb) * sizeof(float);
}

// ---- slow head, stage 1: repetition penalty + per-chunk maximum -------------------------------------------------------------------
struct BHeadArgs {
  bf16 *logits;              // [ncols][V] raw in, penalised out (in place)
  bf16 *logits_raw;          // optional copy of the raw logits (tests) or null
  float *cmax;               // [ncols][nchunk]
  int V, nchunk, n_rows_tok;
  DAState *st;               // [ncols]
};
// grid (nchunk, ncols), 512 threads
__global__ void __launch_bounds__(512) b_head_stats_kernel(const BHeadArgs a) {
  __shared__ float scratch[80];
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y;
  const DAState *st = a.st + n;
  bf16 *lg = a.logits + (size_t)n * a.V;
  const int chunk = (a.V + a.nchunk - 1) / a.nchunk;
  const int i0 = blockIdx.x * chunk, i1 = min(a.V, i0 + chunk);
  const float rp_bf = eff_rep_penalty(st);
  const int use_pen = st->use_penalty;
  __shared__ int s_pen[DA_MAX_ROWS];
  if (threadIdx.x < DA_MAX_ROWS) s_pen[threadIdx.x] = (use_pen && threadIdx.x < a.n_rows_tok) ? st->win[threadIdx.x * DA_WIN] : -1;   // previous_tokens[:, 0]
  __syncthreads();
  float m = -INFINITY;
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    float z = bf2f(lg[i]);
    if (a.logits_raw) a.logits_raw[(size_t)n * a.V + i] = f2bf(z);
    bool hit = false;
#pragma unroll
    for (int r = 0; r < DA_MAX_ROWS; ++r) hit |= (s_pen[r] == i);
    if (hit) { z = penalise(z, rp_bf); lg[i] = f2bf(z); }
    m = fmaxf(m, z);
  }
  m = block_max(m, scratch);
  if (threadIdx.x == 0) a.cmax[(size_t)n * a.nchunk + blockIdx.x] = m;
}

// ---- slow head, stage 2: exact softmax statistics, candidates, sampling in the last CTA of the column ----------------------------
struct BSelectArgs {
  const bf16 *logits; const float *cmax; int V, nchunk; float delta;
  unsigned long long *cand;  // [ncols][DA_CAND_CAP]
  const bf16 *fast_emb; bf16 *fast_x;      // fast_x: [ncols][fast_dim]
  int fast_dim, codebook_size, sem_begin;
  DAState *st;
};
// grid (nchunk, ncols), 512 threads; dynamic smem: scr[192] u64 | scr64[34] | scrf[80]
__global__ void __launch_bounds__(512, 1) b_select_kernel(const BSelectArgs a) {
  extern __shared__ __align__(16) unsigned char smraw_bs[];
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y;
  DAState *st = a.st + n;
  unsigned long long *scr = reinterpret_cast<unsigned long long *>(smraw_bs);
  unsigned long long *scr64 = scr + 192;
  float *scrf = reinterpret_cast<float *>(scr64 + 34);
  __shared__ float s_m;
  if (threadIdx.x < 32) {
    float m = -INFINITY;
    for (int i = threadIdx.x; i < a.nchunk; i += 32) m = fmaxf(m, a.cmax[(size_t)n * a.nchunk + i]);
    m = warp_max(m);
    if (threadIdx.x == 0) s_m = m;
  }
  __syncthreads();
  const float m = s_m, thr = m - a.delta;
  const uint16_t *lb = reinterpret_cast<const uint16_t *>(a.logits + (size_t)n * a.V);
  unsigned long long *cand = a.cand + (size_t)n * DA_CAND_CAP;
  const int chunk = (a.V + a.nchunk - 1) / a.nchunk;
  const int i0 = blockIdx.x * chunk, i1 = min(a.V, i0 + chunk);
  const int lane = threadIdx.x & 31;
  unsigned long long es = 0ull;
  for (int base = i0; base < i1; base += blockDim.x) {
    const int i = base + threadIdx.x;
    uint16_t b = 0; bool c = false;
    if (i < i1) { b = lb[i]; const float z = bits2f(b); c = z >= thr; es += (unsigned long long)(expf(z - m) * DA_FIX2_SCALE); }
    const unsigned mask = __ballot_sync(0xffffffffu, c);
    if (mask) {
      unsigned basei = 0;
      if (lane == 0) basei = atomicAdd(&st->n_cand, (unsigned)__popc(mask));
      basei = __shfl_sync(0xffffffffu, basei, 0);
      if (c) {
        const unsigned slot = basei + __popc(mask & ((1u << lane) - 1));
        if (slot < DA_CAND_CAP) cand[slot] = make_sortkey(b, (uint32_t)i);
      }
    }
  }
  { int par = 0; Red r = {es, 0, -1}; r = block_reduce(r, scr, par); if (threadIdx.x == 0) atomicAdd(&st->s_fix, r.s); __syncthreads(); }
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->sel_ticket, 1u) == (unsigned)a.nchunk - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  SampleParams sp;
  sp.m = m; sp.S = __ull2float_rn(*((volatile unsigned long long *)&st->s_fix)) * (1.0f / DA_FIX2_SCALE);
  sp.T_bf = eff_temperature(st);
  sp.c_max = cmax_from_top_p(st->top_p);
  const unsigned n_cand = *((volatile unsigned *)&st->n_cand);
  uint32_t idx = 0xFFFFFFFFu;
  if (n_cand >= 1 && n_cand <= DA_CAND_CAP) {
    uint32_t key[16], ix[16], valid = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const unsigned e = threadIdx.x + i * blockDim.x;
      key[i] = 0; ix[i] = 0;
      if (e < n_cand) { const unsigned long long k = __ldcg(cand + e); key[i] = 0xFFFFu - (uint32_t)(k >> 32); ix[i] = (uint32_t)k; valid |= 1u << i; }
    }
    idx = sample_items<16>(key, ix, valid, (uint32_t)a.V, (int)n_cand == a.V, sp, noise_src(st), 0u, 0ll, &st->nucleus[0], scr);
    __syncthreads();
  }
  if (idx == 0xFFFFFFFFu) idx = sample_fallback(a.logits + (size_t)n * a.V, a.V, sp, noise_src(st), 0u, 0ll, &st->nucleus[0], scr64, scrf);
  int cb0 = (int)idx - a.sem_begin; if (cb0 < 0) cb0 = 0;                   // inference.py:123-126
  if (cb0 >= a.codebook_size) { cb0 = a.codebook_size - 1; if (threadIdx.x == 0) st->err = 3; }
  for (int d = threadIdx.x; d < a.fast_dim; d += blockDim.x) a.fast_x[(size_t)n * a.fast_dim + d] = a.fast_emb[(size_t)cb0 * a.fast_dim + d];
  if (threadIdx.x == 0) { st->tok_out[0] = (int)idx; st->tok_out[1] = cb0; st->n_cand = 0; st->sel_ticket = 0; st->s_fix = 0ull; }
}

static inline size_t b_fast_sample_smem() { return 256 * 8 + 80 * 4 + (4 * 512 + 2 * 256 + 8) * 4 + 256 * 8 + 4 * 256 * 4 + 64; }
// ---- fast heads: penalty + sampling of one codebook per column; the last head ends the step --------------------------------------
struct BFastSampleArgs {
  const bf16 *logits;        // [ncols][fv] raw
  bf16 *logits_raw;          // optional [ncols][ncb - 1][fv] copy (tests) or null
  int fv, head, ncb, last_head;
  long long noise_off;
  const bf16 *fast_emb; bf16 *fast_x; int fast_dim, codebook_size;
  int *seq; long long seq_slot_stride; int seq_stride, im_end_id, n_rows_tok;
  DAState *st;
};
// grid (ncols), 256 threads; dynamic smem: b_fast_sample_smem()
template <class B> __device__ __forceinline__ void b_fast_sample_body(const BFastSampleArgs &a, int bx, int by, int bz, int gy, unsigned char *dsm) {
  (void)bx; (void)by; (void)bz; (void)gy; (void)dsm;
  unsigned char *smraw_fs = dsm;
  const int n = bx;
  DAState *st = a.st + n;
  unsigned long long *scr = reinterpret_cast<unsigned long long *>(smraw_fs);      // 256 u64: block_reduce buffers + the binned sampler's warp totals
  float *scrf = reinterpret_cast<float *>(scr + 256);                               // 80 floats
  uint32_t *smb = reinterpret_cast<uint32_t *>(scrf + 80);                          // binned sampler: bins, cut list, flags (sampler.cuh)
  __shared__ int s_pen[DA_WIN];
  const bf16 *lg = a.logits + (size_t)n * a.fv;
  const float rp_bf = eff_rep_penalty(st);
  const int use_pen = st->use_penalty;
  if (threadIdx.x < DA_WIN) s_pen[threadIdx.x] = use_pen ? st->win[(a.head + 1) * DA_WIN + threadIdx.x] : -1;      // previous_tokens[k+1]
  B::sync();
  uint32_t it4[DA_FAST_IPT];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < DA_FAST_IPT; ++i) {
    const int e = threadIdx.x * DA_FAST_IPT + i;
    it4[i] = 0xFFFFFFFFu;
    if (e < a.fv) {
      float z = bf2f(lg[e]);
      if (a.logits_raw) a.logits_raw[((size_t)n * (a.ncb - 1) + (a.head - 1)) * a.fv + e] = f2bf(z);
      bool hit = false;
#pragma unroll
      for (int c = 0; c < DA_WIN; ++c) hit |= (s_pen[c] == e);
      if (hit) z = penalise(z, rp_bf);
      it4[i] = ((0xFFFFu - bf16_key(f2bits(z))) << 16) | (uint32_t)e;      // ascending = (logit desc, index asc)
      mx = fmaxf(mx, z);
    }
  }
  SampleParams sp;
  sp.m = block_max<B>(mx, scrf);
  {
    Red es = {0ull, 0, -1}; int par = 0;
#pragma unroll
    for (int i = 0; i < DA_FAST_IPT; ++i) if (it4[i] != 0xFFFFFFFFu) es.s += (unsigned long long)(expf(bits2f(key_bf16(0xFFFFu - (it4[i] >> 16))) - sp.m) * DA_FIX2_SCALE);
    sp.S = __ull2float_rn(block_reduce<B>(es, scr, par).s) * (1.0f / DA_FIX2_SCALE);
    B::sync();
  }
  sp.T_bf = eff_temperature(st);
  sp.c_max = cmax_from_top_p(st->top_p);
  // the binned exact nucleus search (sampler.cuh: same sums, same token as the bisection / sorting samplers) -- 5 block-wide steps
  // instead of ~35 bisection rounds
  uint32_t tok = sample_binned<DA_FAST_IPT, 256, B>(it4, (uint32_t)a.fv, true, nullptr, sp, noise_src(st), (uint32_t)a.head, a.noise_off, &st->nucleus[a.head], smb, scr);
  if (tok >= (uint32_t)a.codebook_size) { tok = a.codebook_size - 1; if (threadIdx.x == 0) st->err = 3; }
  if (!a.last_head) for (int d = threadIdx.x; d < a.fast_dim; d += B::nthreads()) a.fast_x[(size_t)n * a.fast_dim + d] = a.fast_emb[(size_t)tok * a.fast_dim + d];
  B::sync();
  if (threadIdx.x == 0) {
    st->tok_out[a.head + 1] = (int)tok;
    if (a.last_head && !st->done) {
      GemvArgs g; g.st = st; g.seq = a.seq + (size_t)n * a.seq_slot_stride; g.seq_stride = a.seq_stride; g.im_end_id = a.im_end_id; g.n_rows_tok = a.n_rows_tok; g.park_on_done = 1;
      finish_step(g);
    }
  }
}
__global__ void __launch_bounds__(256) b_fast_sample_kernel(const BFastSampleArgs a) {
  extern __shared__ __align__(128) unsigned char dsm_b_fast_sample_body[];
  pdl_launch_dependents();
  pdl_wait();
  b_fast_sample_body<BlockAll>(a, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.y, dsm_b_fast_sample_body);
}

}  // namespace da
