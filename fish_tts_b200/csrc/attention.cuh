// attention.cuh -- slow-AR GQA attention for one query position: split-KV flash-decode.
//
// Replaces Attention.forward llama.py:242-282 for seqlen 1 under SDPBackend.MATH (inference.py:193):
//   q/k nn.RMSNorm (:246-248) -> RoPE (:250-251, 606-618) -> KVCache.update (:142-149, fused: the
//   rotated k and raw v of this position are written straight into the cache) -> repeat_interleave
//   (:258-259, never materialised: a CTA serves all G query heads of one kv head) -> SDPA (:270-274).
// torch's math SDPA upcasts bf16 q/k/v to fp32, scales BOTH q and k by sqrt(scale), takes an fp32
// softmax over the row and an fp32 P@V, and rounds to bf16 once at the end -- we keep exactly those
// rounding points and attend over the valid prefix [0, pos] only (masked columns are exp(-inf) = 0).
//
// Grid (nsplit_max, n_kv_heads).  Each CTA walks its share of the prefix in 64-position tiles; K and
// V tiles arrive in shared memory through the bulk-copy engine (cp.async.bulk + mbarrier, the 1-D
// TMA path: SASS UBLKCP), double buffered.  Partials (m, l, o) go to global memory; the last CTA of
// a kv head to arrive merges them in split order (deterministic) and writes y as bf16.
#pragma once
#include "common.cuh"
#include "gemv.cuh"

namespace da {

#define DA_TILE 64
#define DA_ATTN_THREADS 256
#define DA_MAX_G 8

struct AttnArgs {
  const bf16 *qkv;        // [(nh + 2 nkv) * hd]
  bf16 *kc, *vc;          // [nkv][S][hd]
  const bf16 *rope;       // [S][hd/2][2]
  const bf16 *qn, *kn;    // qk-norm weights or null
  int nh, nkv, hd, S;
  float eps, sf;          // sf = sqrt(1/sqrt(hd)) as float (what math SDPA multiplies q and k by)
  float *part_o;          // [nkv][nsplit_max][G][hd]
  float *part_ml;         // [nkv][nsplit_max][G][2]
  bf16 *y;                // [nh * hd]
  int nsplit_max;
  DAState *st;
  Timeline tl;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_g2s_nohint(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// bounded wait: a lost transaction raises the device fault flag instead of hanging the GPU
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  for (int it = 0; it < (1 << 22) && !done; ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
  return done != 0;
}

// shared memory: q [G][hd] f32 | knew[hd] vnew[hd] f32 | sc [G][TILE] f32 | red 80 f32 | bars 2xu64 | kbuf[2] vbuf[2] bf16 tiles
__global__ void __launch_bounds__(DA_ATTN_THREADS, 2) attn_slow_kernel(const AttnArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  DAState *st = a.st;
  tl_stamp(a.tl, 0);
  pdl_launch_dependents();
  // st->pos / st->done only change in the LAST kernel of a graph launch, and graph launches serialise, so they
  // may be read before the dependency wait; that lets the first K/V tile stream in while the wqkv GEMV finishes
  if (st->done) { pdl_wait(); return; }
  const int g = blockIdx.y, split = blockIdx.x;
  const int G = a.nh / a.nkv, hd = a.hd;
  const int pos = st->pos, L = pos + 1;
  const int n_tiles = (L + DA_TILE - 1) / DA_TILE;
  const int nsplit = min(a.nsplit_max, n_tiles);
  const int tps = (n_tiles + nsplit - 1) / nsplit;
  const int nsplit_eff = (n_tiles + tps - 1) / tps;
  if (split >= nsplit_eff) { pdl_wait(); return; }
  const int t0 = split * tps, t1 = min(n_tiles, t0 + tps);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = DA_ATTN_THREADS / 32;

  float *q = reinterpret_cast<float *>(smraw);
  float *knew = q + G * hd, *vnew = knew + hd;
  float *sc = vnew + hd;
  float *red = sc + G * DA_TILE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(red + 80);
  size_t off = (size_t)((unsigned char *)(bars + 2) - smraw);
  off = (off + 127) & ~(size_t)127;
  bf16 *kbuf = reinterpret_cast<bf16 *>(smraw + off);
  bf16 *vbuf = kbuf + 2 * DA_TILE * hd;

  const bool owns_new = (pos / DA_TILE) >= t0 && (pos / DA_TILE) < t1;
  const uint64_t pol = policy_evict_first();

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // rows of tile t that come from the cache (everything but the position being written now)
  auto issue = [&](int t, int buf) {
    int r0 = t * DA_TILE, r1 = min(L, r0 + DA_TILE);
    int n_old = min(r1, pos) - r0;   // rows < pos
    if (n_old > 0) {
      uint32_t bytes = (uint32_t)n_old * hd * sizeof(bf16);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bars[buf], 2 * bytes);
      bulk_g2s(kbuf + (size_t)buf * DA_TILE * hd, a.kc + ((size_t)g * a.S + r0) * hd, bytes, &bars[buf], pol);
      bulk_g2s(vbuf + (size_t)buf * DA_TILE * hd, a.vc + ((size_t)g * a.S + r0) * hd, bytes, &bars[buf], pol);
    }
  };
  if (threadIdx.x == 0) issue(t0, 0);
  pdl_wait();
  tl_stamp(a.tl, 1);

  // q heads of this group (and, in the split that owns `pos`, the new k / v row)
  const int qd = a.nh * hd, kd = a.nkv * hd;
  {   // one 16-byte chunk per thread: G*hd/8 chunks of q, then hd/8 of the new k and of the new v
    const int nq = (G * hd) >> 3, nk = owns_new ? (hd >> 3) : 0;
    const int c = threadIdx.x;
    if (c < nq + 2 * nk) {
      const bf16 *src; float *dst;
      if (c < nq) { src = a.qkv + (size_t)g * G * hd + (size_t)c * 8; dst = q + c * 8; }
      else if (c < nq + nk) { src = a.qkv + qd + (size_t)g * hd + (size_t)(c - nq) * 8; dst = knew + (c - nq) * 8; }
      else { src = a.qkv + qd + kd + (size_t)g * hd + (size_t)(c - nq - nk) * 8; dst = vnew + (c - nq - nk) * 8; }
      float t[8]; unpack8(*reinterpret_cast<const uint4 *>(src), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = t[j];
    }
  }
  __syncthreads();
  const bf16 *rope_row = a.rope + (size_t)pos * hd;
  for (int h = w; h < G + (owns_new ? 1 : 0); h += nw) {
    if (h < G) head_norm_rope(q + h * hd, hd, a.qn, a.eps, rope_row, lane);
    else head_norm_rope(knew, hd, a.kn, a.eps, rope_row, lane);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < G * hd; e += DA_ATTN_THREADS) q[e] = __fmul_rn(q[e], a.sf);   // q * sqrt(scale), fp32
  if (owns_new)
    for (int e = threadIdx.x; e < hd; e += DA_ATTN_THREADS) {   // KVCache.update
      a.kc[((size_t)g * a.S + pos) * hd + e] = f2bf(knew[e]);
      a.vc[((size_t)g * a.S + pos) * hd + e] = f2bf(vnew[e]);
    }
  __syncthreads();

  // running softmax state: thread `e` owns output element (h, d) = (e / hd, e % hd), e < G*hd
  const int n_own = (G * hd + DA_ATTN_THREADS - 1) / DA_ATTN_THREADS;
  float o_acc[DA_MAX_G * 128 / DA_ATTN_THREADS];
#pragma unroll
  for (int i = 0; i < DA_MAX_G * 128 / DA_ATTN_THREADS; ++i) o_acc[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;   // meaningful in warp h < G (lane-uniform)
  __shared__ float s_m[DA_MAX_G], s_scale[DA_MAX_G], s_l[DA_MAX_G];
  if (threadIdx.x < DA_MAX_G) { s_m[threadIdx.x] = -INFINITY; s_l[threadIdx.x] = 0.f; }

  const int lpr = hd / 8;            // lanes per position row (16-byte pieces)
  const int rpw = 32 / lpr;          // rows per warp pass
  uint32_t phase[2] = {0u, 0u};
  bool ok = true;

  for (int t = t0; t < t1; ++t) {
    const int buf = (t - t0) & 1;
    const int r0 = t * DA_TILE, r1 = min(L, r0 + DA_TILE), nrow = r1 - r0;
    // prefetch the next tile into the other buffer
    if (threadIdx.x == 0 && t + 1 < t1) issue(t + 1, buf ^ 1);
    const int n_old = min(r1, pos) - r0;
    if (n_old > 0) { ok = mbar_wait(&bars[buf], phase[buf]) && ok; phase[buf] ^= 1u; }
    bf16 *kt = kbuf + (size_t)buf * DA_TILE * hd, *vt = vbuf + (size_t)buf * DA_TILE * hd;
    if (pos >= r0 && pos < r1) {   // drop the new row into the tile
      for (int e = threadIdx.x; e < hd; e += DA_ATTN_THREADS) {
        kt[(size_t)(pos - r0) * hd + e] = f2bf(knew[e]);
        vt[(size_t)(pos - r0) * hd + e] = f2bf(vnew[e]);
      }
      __syncthreads();
    }
    // scores: sc[h][j] = sum_d q_s[h][d] * (k[j][d] * sf)
    for (int jb = w * rpw; jb < nrow; jb += nw * rpw) {
      const int j = jb + lane / lpr, piece = lane % lpr;
      float part[DA_MAX_G];
#pragma unroll
      for (int h = 0; h < DA_MAX_G; ++h) part[h] = 0.f;
      if (j < nrow) {
        uint4 kv = *reinterpret_cast<const uint4 *>(kt + (size_t)j * hd + piece * 8);
        float kf[8]; unpack8(kv, kf);
#pragma unroll
        for (int d = 0; d < 8; ++d) kf[d] = __fmul_rn(kf[d], a.sf);
#pragma unroll
        for (int h = 0; h < DA_MAX_G; ++h) {
          if (h < G) {
            const float *qq = q + h * hd + piece * 8;
#pragma unroll
            for (int d = 0; d < 8; ++d) part[h] = fmaf(qq[d], kf[d], part[h]);
          }
        }
      }
#pragma unroll
      for (int h = 0; h < DA_MAX_G; ++h) {
        if (h < G) {
          float v = part[h];
          for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (piece == 0 && j < nrow) sc[h * DA_TILE + j] = v;
        }
      }
    }
    __syncthreads();
    // online softmax bookkeeping: warp h handles head h
    if (w < G) {
      float mx = -INFINITY;
      for (int j = lane; j < nrow; j += 32) mx = fmaxf(mx, sc[w * DA_TILE + j]);
      mx = warp_max(mx);
      float m_new = fmaxf(m_run, mx);
      float ps = 0.f;
      for (int j = lane; j < nrow; j += 32) { float p = expf(sc[w * DA_TILE + j] - m_new); sc[w * DA_TILE + j] = p; ps += p; }
      ps = warp_sum(ps);
      float scale = expf(m_run - m_new);     // exp(-inf) = 0 on the first tile
      l_run = l_run * scale + ps; m_run = m_new;
      if (lane == 0) { s_scale[w] = scale; s_m[w] = m_run; s_l[w] = l_run; }
    }
    __syncthreads();
    // o = o * scale + P @ V
#pragma unroll
    for (int i = 0; i < DA_MAX_G * 128 / DA_ATTN_THREADS; ++i) {
      int e = threadIdx.x + i * DA_ATTN_THREADS;
      if (i < n_own && e < G * hd) {
        int h = e / hd, d = e - h * hd;
        float acc = o_acc[i] * s_scale[h];
        const float *pp = sc + h * DA_TILE;
        for (int j = 0; j < nrow; ++j) acc = fmaf(pp[j], bf2f(vt[(size_t)j * hd + d]), acc);
        o_acc[i] = acc;
      }
    }
    __syncthreads();
  }
  if (!ok && threadIdx.x == 0) st->err = 2;
  tl_stamp(a.tl, 2);

  // partials out
  float *po = a.part_o + (((size_t)g * a.nsplit_max + split) * G) * hd;
  float *pml = a.part_ml + (((size_t)g * a.nsplit_max + split) * G) * 2;
#pragma unroll
  for (int i = 0; i < DA_MAX_G * 128 / DA_ATTN_THREADS; ++i) {
    int e = threadIdx.x + i * DA_ATTN_THREADS;
    if (i < n_own && e < G * hd) po[e] = o_acc[i];
  }
  if (threadIdx.x < G) { pml[threadIdx.x * 2] = s_m[threadIdx.x]; pml[threadIdx.x * 2 + 1] = s_l[threadIdx.x]; }
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->attn_ticket[g], 1u) == (unsigned)nsplit_eff - 1);
  __syncthreads();
  tl_stamp(a.tl, 3);
  if (!s_last) return;
  __threadfence();
  // merge in split order
  for (int e = threadIdx.x; e < G * hd; e += DA_ATTN_THREADS) {
    int h = e / hd;
    float m = -INFINITY;
    for (int s = 0; s < nsplit_eff; ++s)
      m = fmaxf(m, __ldcg(a.part_ml + (((size_t)g * a.nsplit_max + s) * G + h) * 2));
    float l = 0.f, o = 0.f;
    for (int s = 0; s < nsplit_eff; ++s) {
      const float *ml = a.part_ml + (((size_t)g * a.nsplit_max + s) * G + h) * 2;
      float sc_s = expf(__ldcg(ml) - m);
      l = fmaf(__ldcg(ml + 1), sc_s, l);
      o = fmaf(__ldcg(a.part_o + (((size_t)g * a.nsplit_max + s) * G) * hd + e), sc_s, o);
    }
    a.y[(size_t)g * G * hd + e] = f2bf(o / l);
  }
  if (threadIdx.x == 0) st->attn_ticket[g] = 0;
}

static inline size_t attn_smem_bytes(int G, int hd) {
  size_t f = ((size_t)G * hd + 2 * hd + (size_t)G * DA_TILE + 80) * sizeof(float) + 2 * sizeof(uint64_t);
  f = (f + 127) & ~(size_t)127;
  return f + 128 + 4 * (size_t)DA_TILE * hd * sizeof(bf16);
}

}  // namespace da
