// fast_ar.cuh -- the fast-AR codebook loop as ONE persistent, cooperative kernel.
//
// Replaces inference.py:115-149 (forward_generate_fast x num_codebooks interleaved with sample()) -- 170 dependent
// phases per token on s1-mini (10 passes x 4 layers x {wqkv, attention+wo, w1/w3, w2} + 9 heads + 9 samplers), each
// moving only ~25 MB/148 per SM.  As separate kernels every phase pays a launch boundary (3-4 us measured, see
// profiles/); here all SMs stay resident (one CTA per SM, cooperative launch guarantees co-residency) and phases
// hand over through memory alone:
//   * every activation element travels as one 32-bit UNIT = bf16 value in the high half | 16-bit phase tag in the low
//     half, written with a single 4-byte store.  A consumer polls the units it needs until the tag matches the
//     producing phase -- data and "ready" flag arrive atomically, so no fence, no atomic counter and no grid barrier
//     sits between two phases (the idea of NCCL's LL protocol, applied on-chip through L2).  Tags advance with a
//     per-engine phase counter, so consecutive uses of a buffer never share a tag and nothing needs clearing.
//   * weights are not phase-ordered: each warp always has the first 128-bit batch of its NEXT row pair in flight
//     (possibly several phases ahead), so the L2 / HBM stream of the fast stack never waits for a hand-over.
//   * the fast KV cache of the current token lives in shared memory (every CTA recomputes the <=16-head attention
//     of one position, exactly like the multi-kernel path), so no cross-CTA memory ordering is needed for it.
// Row dot products use the same canonical order as gemv.cuh (lane-strided 8-element chunks, fp32 fma chain, butterfly),
// so this kernel and the multi-kernel path agree bit for bit (tests/test_gpu_parity.py::test_fast_ar_kernel_bitexact).
// All spins are bounded: a lost hand-over raises the device fault flag instead of hanging the GPU.
#pragma once
#include "common.cuh"
#include "gemv.cuh"
#include "sampler.cuh"

namespace da {

#define DA_MAX_FAST_LAYERS 8
#define DA_FAR_THREADS 512
#define DA_SPIN_LIMIT (1 << 22)

struct FastLayerW {
  const bf16 *wqkv, *bqkv, *wo, *bo, *qn, *kn, *w13, *w2, *ffn_norm, *attn_norm;
};

struct FastArArgs {
  FastLayerW L[DA_MAX_FAST_LAYERS];
  int n_layer;
  const bf16 *fast_norm, *fast_out, *fast_emb, *rope;
  int dim, nh, nkv, hd, inter, ncb, fv, codebook_size;
  float eps, scale;
  const bf16 *x_slow;       // plain bf16 [dim]: input of pass 0 (the slow stack's un-normalised hidden state)
  const bf16 *fin_plain;    // plain bf16 [dim]: input of pass 1 (embedding of codebook 0, written by the slow-head sampler)
  uint32_t *u_qkv, *u_h, *u_act, *u_x0, *u_x1, *u_fin, *u_logits;   // unit buffers
  bf16 *flogits_raw;        // [(ncb-1)][fv] plain copy before the penalty (introspection)
  bf16 *flogits;            // [(ncb-1)][fv] plain copy after the penalty
  int *seq; int seq_stride, im_end_id, n_rows_tok;
  long long noise_off0;     // vocab_size: offset of fast head 1 inside a step's noise block
  DAState *st;
  Timeline tl;
};

__device__ __forceinline__ uint32_t make_unit(float v, uint32_t tag) { return ((uint32_t)f2bits(v) << 16) | tag; }
__device__ __forceinline__ float unit_val(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint4 ld_poll4(const uint32_t *p) {
  uint4 r;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ uint32_t ld_poll1(const uint32_t *p) {
  uint32_t r;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_unit(uint32_t *p, uint32_t u) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(u) : "memory");
}
__device__ __forceinline__ bool tags_ok(const uint4 &u, uint32_t tag) {
  return ((u.x & 0xFFFFu) == tag) & ((u.y & 0xFFFFu) == tag) & ((u.z & 0xFFFFu) == tag) & ((u.w & 0xFFFFu) == tag);
}
// poll 8 consecutive units (one 8-element chunk) until every tag matches; false on timeout
__device__ __forceinline__ bool poll_chunk(const uint32_t *p, uint32_t tag, float *f) {
  uint4 a, b;
  int it = 0;
  do {
    a = ld_poll4(p); b = ld_poll4(p + 4);
    if (tags_ok(a, tag) & tags_ok(b, tag)) break;
    __nanosleep(100);
  } while (++it < DA_SPIN_LIMIT);
  f[0] = unit_val(a.x); f[1] = unit_val(a.y); f[2] = unit_val(a.z); f[3] = unit_val(a.w);
  f[4] = unit_val(b.x); f[5] = unit_val(b.y); f[6] = unit_val(b.z); f[7] = unit_val(b.w);
  return it < DA_SPIN_LIMIT;
}

// Polling etiquette: unthrottled, all-thread polling saturates L2 (148 SMs x hundreds of relaxed loads in flight) and
// starves the weight stream.  So ONE thread per CTA first watches a single sentinel unit -- the last element of the
// vector, produced late in its phase -- with back-off; only then does every thread read its own chunk (tags usually
// match on the first try).
__device__ __forceinline__ bool wait_sentinel(const uint32_t *p, uint32_t tag) {
  int it = 0;
  while ((ld_poll1(p) & 0xFFFFu) != tag) { if (++it >= DA_SPIN_LIMIT) return false; __nanosleep(60); }
  return true;
}

// ---- the static schedule -----------------------------------------------------------------------------
// linear phase index ph -> (pass p, slot s): per pass n_layer*4 GEMV phases [qkv, wo, w13, w2] then (p >= 1) a head phase.
// Sampling is not a phase of its own: CTA 0 samples right after polling the head's logits.
enum { FK_QKV = 0, FK_WO = 1, FK_W13 = 2, FK_W2 = 3, FK_HEAD = 4 };
struct PhaseInfo { int pass, layer, kind; const bf16 *W; const bf16 *bias; int rows, K; };

__device__ __forceinline__ int far_phases_per_pass0(const FastArArgs &a) { return a.n_layer * 4; }
__device__ __forceinline__ int far_num_phases(const FastArArgs &a) { return a.ncb * a.n_layer * 4 + (a.ncb - 1); }
__device__ __forceinline__ PhaseInfo far_phase(const FastArArgs &a, int ph) {
  PhaseInfo r;
  const int p0 = a.n_layer * 4, pp = p0 + 1;
  int pass, s;
  if (ph < p0) { pass = 0; s = ph; } else { pass = 1 + (ph - p0) / pp; s = (ph - p0) % pp; }
  r.pass = pass;
  if (s == p0) {
    r.kind = FK_HEAD; r.layer = a.n_layer - 1; r.W = a.fast_out; r.bias = nullptr; r.rows = a.fv; r.K = a.dim;
    return r;
  }
  r.layer = s >> 2; r.kind = s & 3;
  const FastLayerW &L = a.L[r.layer];
  const int qd = a.nh * a.hd, kd = a.nkv * a.hd;
  switch (r.kind) {
    case FK_QKV: r.W = L.wqkv; r.bias = L.bqkv; r.rows = qd + 2 * kd; r.K = a.dim; break;
    case FK_WO: r.W = L.wo; r.bias = L.bo; r.rows = a.dim; r.K = qd; break;
    case FK_W13: r.W = L.w13; r.bias = nullptr; r.rows = 2 * a.inter; r.K = a.dim; break;
    default: r.W = L.w2; r.bias = nullptr; r.rows = a.dim; r.K = a.inter; break;
  }
  return r;
}

// shared memory: xs[4096] f32 | scratch 80 f32 | q[nh*hd] f32 | kcur,vcur [nkv*hd] f32 | pr [nh*ncb] f32 |
//                sampler scratch 192 u64 | kv store bf16 [n_layer][ncb][2][nkv*hd]
static inline size_t far_smem_bytes(int n_layer, int nh, int nkv, int hd, int ncb) {
  size_t f = (size_t)(4096 + 80 + nh * hd + 2 * nkv * hd + nh * ncb) * sizeof(float) + 192 * 8;
  f = (f + 15) & ~(size_t)15;
  return f + (size_t)n_layer * ncb * 2 * nkv * hd * sizeof(bf16) + 64;
}

__global__ void __launch_bounds__(DA_FAR_THREADS, 1) fast_ar_kernel(const FastArArgs a) {
  extern __shared__ __align__(16) float smem_far[];
  DAState *st = a.st;
  tl_stamp(a.tl, 0);
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int qd = a.nh * a.hd, kd = a.nkv * a.hd, G = a.nh / a.nkv;
  float *xs = smem_far, *scratch = xs + 4096, *q = scratch + 80, *kcur = q + qd, *vcur = kcur + kd, *pr = vcur + kd;
  unsigned long long *scr = reinterpret_cast<unsigned long long *>(pr + a.nh * a.ncb + ((a.nh * a.ncb) & 1));
  bf16 *kvs = reinterpret_cast<bf16 *>(scr + 192);
  const uint64_t pol = policy_evict_last();
  const int nph = far_num_phases(a);
  const int first = w * gridDim.x + blockIdx.x, stride = nw * gridDim.x;

  // ---- weight cursor: (phase, pair index, batch) of the unit held in `ra` / `rb` ------------------------------
  PairRegs ra, rb;
  int c_ph = 0, c_pi = 0, c_b = 0;        // the unit currently loaded / being consumed
  bool c_valid = false;
  auto my_pairs = [&](const PhaseInfo &pi_) { const int np = (pi_.rows + 1) >> 1; return first < np ? (np - first + stride - 1) / stride : 0; };
  // advance (ph, pi, b) to the next unit of this warp; returns false past the end
  auto next_unit = [&](int &ph, int &pi, int &b) {
    PhaseInfo f = far_phase(a, ph);
    const int nb = ((f.K >> 8) + DA_CH - 1) / DA_CH;
    if (b + 1 < nb) { ++b; return true; }
    b = 0;
    if (pi + 1 < my_pairs(f)) { ++pi; return true; }
    pi = 0;
    for (++ph; ph < nph; ++ph) if (my_pairs(far_phase(a, ph)) > 0) return true;
    return false;
  };
  auto issue = [&](PairRegs &r, int ph, int pi, int b) {
    PhaseInfo f = far_phase(a, ph);
    load_batch(r, f.W, f.K, f.rows, first + pi * stride, b, f.K >> 8, lane, pol);
  };
  // first unit of this warp (weights only: may be fetched before the dependency wait)
  { int ph = 0; while (ph < nph && my_pairs(far_phase(a, ph)) == 0) ++ph; if (ph < nph) { c_ph = ph; c_valid = true; issue(ra, c_ph, 0, 0); } }
  int parity = 0;                          // 0: current unit in ra, next goes to rb

  pdl_wait();
  tl_stamp(a.tl, 1);
  if (*reinterpret_cast<const volatile int *>(&st->done)) return;
  const uint32_t tag_base = *reinterpret_cast<const volatile unsigned int *>(&st->phase_ctr);
  auto tag_of = [&](int ph) { return (uint32_t)((tag_base + (uint32_t)ph) % 65535u) + 1u; };
  const float rp_eff = eff_rep_penalty(st);
  const int use_pen = st->use_penalty;
  bool ok = true;

  for (int ph = 0; ph < nph; ++ph) {
    const PhaseInfo f = far_phase(a, ph);
    const uint32_t tag = tag_of(ph);
    const int nseg = f.K >> 8, nb = (nseg + DA_CH - 1) / DA_CH;
    Timeline tp = a.tl; tp.slot = 192 + ph;
    if (192 + ph < 512) tl_stamp(tp, 0);
    // ---- (A) stage this phase's input vector into xs ------------------------------------------------------------
    // which buffer feeds this phase, and the tag of the phase that wrote it
    const uint32_t *in_u = nullptr; const bf16 *in_plain = nullptr; int in_ph = ph - 1;
    const uint32_t *res_u = nullptr; const bf16 *res_plain = nullptr;
    // the layer input (needed by QKV as x and by WO as residual)
    const uint32_t *lin_u = nullptr; const bf16 *lin_plain = nullptr; int lin_ph = 0;
    if (f.layer == 0 || f.kind == FK_HEAD) {
      if (f.kind != FK_HEAD) {
        if (f.pass == 0) lin_plain = a.x_slow;
        else if (f.pass == 1) lin_plain = a.fin_plain;
        else { lin_u = a.u_fin; lin_ph = a.n_layer * 4 + (f.pass - 2) * (a.n_layer * 4 + 1) + a.n_layer * 4; }   // head phase of pass-1
      }
    } else {
      lin_u = ((f.layer - 1) & 1) ? a.u_x1 : a.u_x0;
      lin_ph = ph - f.kind - 1;            // the w2 phase of the previous layer
    }
    uint32_t *out_u = nullptr;
    const bf16 *norm_w = nullptr;
    switch (f.kind) {
      case FK_QKV: in_u = lin_u; in_plain = lin_plain; in_ph = lin_ph; out_u = a.u_qkv; norm_w = a.L[f.layer].attn_norm; break;
      case FK_WO: in_u = a.u_qkv; res_u = lin_u; res_plain = lin_plain; out_u = a.u_h; break;
      case FK_W13: in_u = a.u_h; out_u = a.u_act; norm_w = a.L[f.layer].ffn_norm; break;
      case FK_W2: in_u = a.u_act; res_u = a.u_h; out_u = (f.layer & 1) ? a.u_x1 : a.u_x0; break;
      default: in_u = ((a.n_layer - 1) & 1) ? a.u_x1 : a.u_x0; out_u = a.u_logits; norm_w = a.fast_norm; break;
    }
    const uint32_t in_tag = tag_of(in_ph);

    if (f.kind == FK_WO) {
      // ---- fast-layer attention for position `pass` (llama.py:246-251, 285-309), recomputed by every CTA -------------
      const int pos = f.pass, P = pos + 1;
      bf16 *kv_l = kvs + (size_t)f.layer * a.ncb * 2 * kd;
      if (threadIdx.x == 0) ok = wait_sentinel(in_u + qd + 2 * kd - 1, in_tag) && ok;
      __syncthreads();
      {
        const int c = threadIdx.x;                 // chunk of 8 units of [q | k | v]
        if (c * 8 < qd + 2 * kd) {
          float t[8];
          ok = poll_chunk(in_u + c * 8, in_tag, t) && ok;
          const int e = c * 8;
          float *dst = e < qd ? q + e : (e < qd + kd ? kcur + (e - qd) : vcur + (e - qd - kd));
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = t[j];
        }
      }
      __syncthreads();
      const bf16 *rope_row = a.rope + (size_t)pos * a.hd;
      for (int h = w; h < a.nh + a.nkv; h += nw) {
        if (h < a.nh) head_norm_rope(q + h * a.hd, a.hd, a.L[f.layer].qn, a.eps, rope_row, lane);
        else head_norm_rope(kcur + (h - a.nh) * a.hd, a.hd, a.L[f.layer].kn, a.eps, rope_row, lane);
      }
      __syncthreads();
      for (int e = threadIdx.x; e < kd; e += blockDim.x) {          // this token's fast KV cache row (shared memory)
        kv_l[((size_t)pos * 2 + 0) * kd + e] = f2bf(kcur[e]);
        kv_l[((size_t)pos * 2 + 1) * kd + e] = f2bf(vcur[e]);
      }
      __syncthreads();
      for (int t = threadIdx.x; t < a.nh * P; t += blockDim.x) {    // scores: bf16(q @ k^T), then bf16(* scale)
        const int h = t / P, j = t - h * P, g = h / G;
        const float *qq = q + h * a.hd;
        const bf16 *kk = kv_l + ((size_t)j * 2 + 0) * kd + g * a.hd;
        float acc = 0.f;
        for (int d = 0; d < a.hd; ++d) acc = fmaf(qq[d], bf2f(kk[d]), acc);
        pr[h * a.ncb + j] = rbf(__fmul_rn(rbf(acc), a.scale));
      }
      __syncthreads();
      float pval = 0.f;
      const int t = threadIdx.x;
      if (t < a.nh * P) {                                              // softmax (fp32, rounded to bf16), one thread per (h, j)
        const int h = t / P, j = t - h * P;
        float m = -INFINITY;
        for (int jj = 0; jj < P; ++jj) m = fmaxf(m, pr[h * a.ncb + jj]);
        float sum = 0.f;
        for (int jj = 0; jj < P; ++jj) sum += expf(pr[h * a.ncb + jj] - m);
        pval = rbf(expf(pr[h * a.ncb + j] - m) / sum);
      }
      __syncthreads();
      if (t < a.nh * P) { const int h = t / P, j = t - h * P; pr[h * a.ncb + j] = pval; }
      __syncthreads();
      for (int c = threadIdx.x; c * 8 < qd; c += blockDim.x) {        // y = bf16(p @ v) -> the wo input
        const int e = c * 8, h = e / a.hd, d = e - h * a.hd, g = h / G;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int jj = 0; jj < P; ++jj) {
          const float pj = pr[h * a.ncb + jj];
          const bf16 *vv = kv_l + ((size_t)jj * 2 + 1) * kd + g * a.hd + d;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(pj, bf2f(vv[j]), acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = rbf(acc[j]);
        store_chunk_xs(xs, c, acc);
      }
    } else {
      // ---- PLAIN / RMSNORM prologue over units (or the plain bf16 vector at the kernel boundary) -----------------------------
      const int c = threadIdx.x;
      const bool mine = c * 8 < f.K;
      float v[8], g[8];
      if (in_u) {
        if (threadIdx.x == 0) ok = wait_sentinel(in_u + f.K - 1, in_tag) && ok;
        __syncthreads();
      }
      if (mine) {
        if (norm_w) unpack8(*reinterpret_cast<const uint4 *>(norm_w + (size_t)c * 8), g);
        if (in_u) ok = poll_chunk(in_u + c * 8, in_tag, v) && ok;
        else unpack8(*reinterpret_cast<const uint4 *>(in_plain + (size_t)c * 8), v);
      }
      if (norm_w) {
        float ss = 0.f;
        if (mine) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
        }
        ss = block_sum(ss, scratch);
        const float inv = rsqrtf(ss * (1.0f / (float)f.K) + a.eps);
        if (mine) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = rbf(__fmul_rn(rbf(__fmul_rn(v[j], inv)), g[j]));
        }
      }
      if (mine) store_chunk_xs(xs, c, v);
    }
    __syncthreads();
    if (ph == 0) tl_stamp(a.tl, 2);
    if (192 + ph < 512) tl_stamp(tp, 1);

    // ---- (B) this warp's row pairs of the phase ----------------------------------------------------------------------
    // LOGITS: penalised ids for this head (previous_tokens[k+1], inference.py:141-145)
    int pen_id = -1;
    if (f.kind == FK_HEAD && use_pen && lane < DA_WIN) pen_id = st->win[(f.pass + 1) * DA_WIN + lane];
    float a0 = 0.f, a1 = 0.f;
    while (c_valid && c_ph == ph) {
      int n_ph = c_ph, n_pi = c_pi, n_b = c_b;
      const bool has_next = next_unit(n_ph, n_pi, n_b);
      if (parity == 0) { if (has_next) issue(rb, n_ph, n_pi, n_b); fma_batch(ra, xs, c_b, nseg, lane, a0, a1); }
      else { if (has_next) issue(ra, n_ph, n_pi, n_b); fma_batch(rb, xs, c_b, nseg, lane, a0, a1); }
      parity ^= 1;
      if (c_b == nb - 1) {
        const int p = first + c_pi * stride;
        const int r0 = 2 * p, r1 = r0 + 1;
        const bool has1 = r1 < f.rows;
        float d0 = warp_sum(a0), d1 = warp_sum(a1);
        a0 = 0.f; a1 = 0.f;
        if (f.bias) { d0 += bf2f(f.bias[r0]); if (has1) d1 += bf2f(f.bias[r1]); }
        if (f.kind == FK_QKV) {
          if (lane == 0) { st_unit(out_u + r0, make_unit(d0, tag)); if (has1) st_unit(out_u + r1, make_unit(d1, tag)); }
        } else if (f.kind == FK_WO || f.kind == FK_W2) {
          if (lane == 0) {
            float q0, q1 = 0.f;
            if (res_u) { q0 = unit_val(ld_poll1(res_u + r0)); if (has1) q1 = unit_val(ld_poll1(res_u + r1)); }
            else { q0 = bf2f(res_plain[r0]); if (has1) q1 = bf2f(res_plain[r1]); }
            st_unit(out_u + r0, make_unit(q0 + rbf(d0), tag));
            if (has1) st_unit(out_u + r1, make_unit(q1 + rbf(d1), tag));
          }
        } else if (f.kind == FK_W13) {
          if (lane == 0) {
            const float gg = rbf(d0), up = rbf(d1);
            const float sg = rbf(gg / (1.0f + expf(-gg)));
            st_unit(out_u + p, make_unit(__fmul_rn(sg, up), tag));
          }
        } else {
          float z0 = rbf(d0), z1 = rbf(d1);
          const size_t lo = (size_t)(f.pass - 1) * a.fv;
          if (lane == 0) { a.flogits_raw[lo + r0] = f2bf(z0); if (has1) a.flogits_raw[lo + r1] = f2bf(z1); }
          const unsigned hit0 = __ballot_sync(0xffffffffu, pen_id == r0), hit1 = __ballot_sync(0xffffffffu, pen_id == r1);
          if (hit0) z0 = penalise(z0, rp_eff);
          if (hit1) z1 = penalise(z1, rp_eff);
          if (lane == 0) {
            a.flogits[lo + r0] = f2bf(z0); st_unit(out_u + r0, make_unit(z0, tag));
            if (has1) { a.flogits[lo + r1] = f2bf(z1); st_unit(out_u + r1, make_unit(z1, tag)); }
          }
        }
      }
      c_ph = n_ph; c_pi = n_pi; c_b = n_b; c_valid = has_next;
    }

    if (192 + ph < 512) tl_stamp(tp, 2);
    // ---- (C) the head's sampler: CTA 0 polls all logits, draws the code, publishes its embedding as pass p+1's input ----------
    if (f.kind == FK_HEAD) {
      if (blockIdx.x == 0) {
        const int V = a.fv;
        if (threadIdx.x == 0) ok = wait_sentinel(a.u_logits + V - 1, tag) && ok;
        __syncthreads();
        uint32_t key[2], idx[2], valid = 0;
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int e = threadIdx.x * 2 + i;
          idx[i] = (uint32_t)e; key[i] = 0;
          if (e < V) {
            uint32_t u; int it = 0;
            do { u = ld_poll1(a.u_logits + e); if ((u & 0xFFFFu) == tag) break; __nanosleep(100); } while (++it < DA_SPIN_LIMIT);
            ok = ok && it < DA_SPIN_LIMIT;
            key[i] = bf16_key((uint16_t)(u >> 16)); valid |= 1u << i; mx = fmaxf(mx, unit_val(u));
          }
        }
        SampleParams sp;
        sp.m = block_max(mx, scratch);
        float es = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) if ((valid >> i) & 1u) es += expf(bits2f(key_bf16(key[i])) - sp.m);
        sp.S = block_sum(es, scratch);
        sp.T_bf = eff_temperature(st);
        sp.c_max = cmax_from_top_p(st->top_p);
        uint32_t tok = sample_items<2>(key, idx, valid, (uint32_t)V, true, sp, st, (uint32_t)f.pass,
                                       a.noise_off0 + (long long)(f.pass - 1) * a.fv, &st->nucleus[f.pass], scr);
        if (tok >= (uint32_t)a.codebook_size) { tok = a.codebook_size - 1; if (threadIdx.x == 0) st->err = 3; }
        if (f.pass < a.ncb - 1)
          for (int d = threadIdx.x; d < a.dim; d += blockDim.x) st_unit(a.u_fin + d, make_unit(bf2f(a.fast_emb[(size_t)tok * a.dim + d]), tag));
        if (threadIdx.x == 0) st->tok_out[f.pass + 1] = (int)tok;
        __syncthreads();
      }
    }
  }

  tl_stamp(a.tl, 3);
  if (!ok && threadIdx.x == 0) st->err = 4;
  if (blockIdx.x == 0) {
    __syncthreads();
    if (threadIdx.x == 0) {
      st->phase_ctr = tag_base + (unsigned)nph;
      GemvArgs g; g.st = st; g.seq = a.seq; g.seq_stride = a.seq_stride; g.im_end_id = a.im_end_id; g.n_rows_tok = a.n_rows_tok;
      finish_step(g);
    }
  }
}

}  // namespace da
