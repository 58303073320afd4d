// misc_kernels.cuh -- the small kernels around the GEMVs: token embedding, the slow-head
// candidate selection + sampler, step-input loading and the noise generator.
#pragma once
#include "common.cuh"
#include "sampler.cuh"
#include "gemv.cuh"

namespace da {

// ---- K1: token + codebook embedding (llama.py:409-429) -------------------------------------------
struct EmbedArgs {
  const bf16 *emb;        // [vocab][dim]
  const bf16 *cb_emb;     // [codebook_size * num_codebooks][dim]
  bf16 *x;                // [dim]
  int dim, vocab, codebook_size, num_codebooks;
  int sem_begin, sem_end, scale_cb;
  float inv_sqrt;         // float(1 / sqrt(num_codebooks + 1)): CUDA `tensor / python_scalar` multiplies by the reciprocal
  float sqrt_c;           // float(sqrt(num_codebooks + 1)): the CPU kernel divides
  DAState *st;
  Timeline tl;
};

__global__ void __launch_bounds__(256) embed_kernel(const EmbedArgs a) {
  DAState *st = a.st;
  tl_stamp(a.tl, 0);
  pdl_launch_dependents();
  pdl_wait();
  tl_stamp(a.tl, 1);
  if (st->done) return;
  int tok = st->tok_in[0];
  if (tok < 0 || tok >= a.vocab) { tok = 0; if (threadIdx.x == 0) st->err = 1; }
  const bool is_sem = tok >= a.sem_begin && tok <= a.sem_end;
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < a.dim; d += gridDim.x * blockDim.x) {
    float vq = 0.f;
    if (is_sem) {
      for (int i = 0; i < a.num_codebooks; ++i) {       // stack(...).sum(dim=1): fp32 accumulate, one rounding
        int c = st->tok_in[i + 1];
        if (c < 0 || c >= a.codebook_size) { c = 0; st->err = 1; }
        vq += bf2f(a.cb_emb[((size_t)c + (size_t)i * a.codebook_size) * a.dim + d]);
      }
      vq = rbf(vq);
    }
    float x = rbf(bf2f(a.emb[(size_t)tok * a.dim + d]) + vq);
    if (a.scale_cb && is_sem) x = st->cpu_sem ? rbf(__fdiv_rn(x, a.sqrt_c)) : rbf(__fmul_rn(x, a.inv_sqrt));
    a.x[d] = f2bf(x);
  }
  tl_stamp(a.tl, 3);
}

// ---- slow head, stage 2: candidate selection on all SMs, sampling in the last CTA -----------------
struct SelectArgs {
  const bf16 *logits;     // [V] penalised
  const float2 *partials; int n_partials;
  int V;
  float delta;            // candidates: z >= max - delta
  unsigned long long *cand;   // [DA_CAND_CAP] global
  const bf16 *fast_emb; bf16 *fast_x; int fast_dim, codebook_size, sem_begin;
  DAState *st;
  Timeline tl;
};

#define DA_SEL_IPT 16   // 512 threads x 16 items = DA_CAND_CAP
// dynamic smem: scr[192] u64 | scr64[34] | scrf[80]
__global__ void __launch_bounds__(512, 1) select_sample_kernel(const SelectArgs a) {
  extern __shared__ __align__(16) unsigned char smraw_sel[];
  DAState *st = a.st;
  tl_stamp(a.tl, 0);
  pdl_launch_dependents();
  pdl_wait();
  tl_stamp(a.tl, 1);
  if (st->done) return;
  unsigned long long *scr = reinterpret_cast<unsigned long long *>(smraw_sel);
  unsigned long long *scr64 = scr + 192;
  float *scrf = reinterpret_cast<float *>(scr64 + 34);
  __shared__ float s_m;
  // the global max from the head kernel's per-CTA partials (exact, order-free)
  if (threadIdx.x < 32) {
    float m = -INFINITY;
    for (int i = threadIdx.x; i < a.n_partials; i += 32) m = fmaxf(m, a.partials[i].x);
    m = warp_max(m);
    if (threadIdx.x == 0) s_m = m;
  }
  __syncthreads();
  const float m = s_m, thr = m - a.delta;
  const uint16_t *lb = reinterpret_cast<const uint16_t *>(a.logits);
  const int chunk = (a.V + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * chunk, i1 = min(a.V, i0 + chunk);
  const int lane = threadIdx.x & 31;
  // S = sum exp(z - m) as 2^-40 fixed point: independent of summation order and of how rows were dealt to CTAs,
  // so every kernel path computes the same S bit for bit
  unsigned long long es = 0ull;
  for (int base = i0; base < i1; base += blockDim.x) {
    int i = base + threadIdx.x;
    uint16_t b = 0; bool c = false;
    if (i < i1) { b = lb[i]; const float z = bits2f(b); c = z >= thr; es += (unsigned long long)(expf(z - m) * DA_FIX2_SCALE); }
    unsigned mask = __ballot_sync(0xffffffffu, c);
    if (mask) {
      unsigned basei = 0;
      if (lane == 0) basei = atomicAdd(&st->n_cand, (unsigned)__popc(mask));
      basei = __shfl_sync(0xffffffffu, basei, 0);
      if (c) {
        unsigned slot = basei + __popc(mask & ((1u << lane) - 1));
        if (slot < DA_CAND_CAP) a.cand[slot] = make_sortkey(b, (uint32_t)i);
      }
    }
  }
  { int par = 0; Red r = {es, 0, -1}; r = block_reduce(r, scr, par); if (threadIdx.x == 0) atomicAdd(&st->s_fix, r.s); __syncthreads(); }
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->sel_ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  tl_stamp(a.tl, 2);
  if (!s_last) return;
  __threadfence();

  SampleParams sp;
  sp.m = m; sp.S = __ull2float_rn(*((volatile unsigned long long *)&st->s_fix)) * (1.0f / DA_FIX2_SCALE);
  sp.T_bf = eff_temperature(st);
  sp.c_max = cmax_from_top_p(st->top_p);
  const unsigned n_cand = *((volatile unsigned *)&st->n_cand);
  uint32_t idx = 0xFFFFFFFFu;
  if (n_cand >= 1 && n_cand <= DA_CAND_CAP) {
    uint32_t key[DA_SEL_IPT], ix[DA_SEL_IPT], valid = 0;
#pragma unroll
    for (int i = 0; i < DA_SEL_IPT; ++i) {
      const unsigned e = threadIdx.x + i * blockDim.x;   // coalesced reads of the (unordered) candidate list
      key[i] = 0; ix[i] = 0;
      if (e < n_cand) { const unsigned long long k = __ldcg(a.cand + e); key[i] = 0xFFFFu - (uint32_t)(k >> 32); ix[i] = (uint32_t)k; valid |= 1u << i; }
    }
    idx = sample_items<DA_SEL_IPT>(key, ix, valid, (uint32_t)a.V, (int)n_cand == a.V, sp, noise_src(st), 0u, 0ll, &st->nucleus[0], scr);
    __syncthreads();
  }
  if (idx == 0xFFFFFFFFu) idx = sample_fallback(a.logits, a.V, sp, noise_src(st), 0u, 0ll, &st->nucleus[0], scr64, scrf);
  // inference.py:123-126: first codebook = semantic id - semantic_begin (clamped at 0); next input = its fast embedding
  int cb0 = (int)idx - a.sem_begin; if (cb0 < 0) cb0 = 0;
  if (cb0 >= a.codebook_size) { cb0 = a.codebook_size - 1; if (threadIdx.x == 0) st->err = 3; }
  for (int d = threadIdx.x; d < a.fast_dim; d += blockDim.x) a.fast_x[d] = a.fast_emb[(size_t)cb0 * a.fast_dim + d];
  if (threadIdx.x == 0) {
    st->tok_out[0] = (int)idx; st->tok_out[1] = cb0;
    st->n_cand = 0; st->sel_ticket = 0; st->s_fix = 0ull;
    if (a.tl.buf) { unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); a.tl.buf[a.tl.slot * 8 + 3] = g; }
  }
}

// ---- step-mode input loading: what the reference passes to decode_one_token (inference.py:194-204)
struct LoadStepArgs {
  const int *x; const int *input_pos; const int *prev; long long prev_stride;
  const float *temperature, *top_p, *rep_penalty; const bf16 *noise;
  int n_rows;
  DAState *st;
};
__global__ void load_step_kernel(const LoadStepArgs a) {
  DAState *st = a.st;
  int t = threadIdx.x;
  if (t < a.n_rows) st->tok_in[t] = a.x[t];
  if (a.prev) for (int i = t; i < a.n_rows * DA_WIN; i += blockDim.x) st->win[i] = a.prev[(size_t)(i / DA_WIN) * a.prev_stride + (i % DA_WIN)];
  if (t == 0) {
    st->pos = a.input_pos[0];
    st->temperature = a.temperature[0]; st->top_p = a.top_p[0]; st->rep_penalty = a.rep_penalty[0];
    st->use_penalty = a.prev ? 1 : 0;
    st->noise = a.noise; st->loop_mode = 0; st->done = 0; st->err = 0;
    st->n_cand = 0; st->sel_ticket = 0; st->fast_ticket = 0; st->head_ticket = 0; st->s_fix = 0ull;
    for (int g = 0; g < DA_MAX_KV_HEADS; ++g) st->attn_ticket[g] = 0;
  }
}
__global__ void store_step_kernel(const DAState *st, int *out, int n_rows, int *err_sticky) {
  if (threadIdx.x < n_rows) out[threadIdx.x] = st->tok_out[threadIdx.x];
  if (threadIdx.x == 0) {
    const_cast<DAState *>(st)->step_ctr += 1;
    if (st->err && err_sticky) *err_sticky = st->err;      // mapped host word: the next dualar_step call reports it
  }
}

// ---- loop-mode prefill bookkeeping -----------------------------------------------------------------
// loads column `pos` of the prompt as the step input; on the last prompt position arms the sampler
struct PrefillColArgs { const int *seq; int seq_stride; int n_rows; DAState *st; };
__global__ void prefill_col_kernel(const PrefillColArgs a, int advance) {
  DAState *st = a.st;
  pdl_launch_dependents();
  pdl_wait();
  if (advance) { if (threadIdx.x == 0) st->pos += 1; }
  __syncthreads();
  int pos = st->pos;
  if (threadIdx.x < a.n_rows) st->tok_in[threadIdx.x] = a.seq[(size_t)threadIdx.x * a.seq_stride + pos];
}

// ---- test hook: run the sampler alone on caller-supplied logits (dualar_debug_sample) ------------------
// slow head: penalty + per-CTA softmax partials, exactly what the LOGITS epilogue leaves behind
__global__ void __launch_bounds__(512) debug_stats_kernel(const bf16 *in, bf16 *out, float2 *partials, int V, int n_rows_tok, DAState *st) {
  __shared__ float scratch[80];
  const int chunk = (V + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * chunk, i1 = min(V, i0 + chunk);
  const float rp_bf = eff_rep_penalty(st);
  float m = -INFINITY;
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    float z = bf2f(in[i]);
    if (st->use_penalty) for (int r = 0; r < n_rows_tok; ++r) if (st->win[r * DA_WIN] == i) { z = penalise(z, rp_bf); break; }
    out[i] = f2bf(z);
    m = fmaxf(m, z);
  }
  m = block_max(m, scratch);
  float l = 0.f;
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) l += expf(bf2f(out[i]) - m);
  l = block_sum(l, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = make_float2(m, l);
}

__global__ void fill_noise_kernel(bf16 *out, long long n, unsigned long long seed, unsigned step, unsigned head) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = f2bf(exp1_noise(seed, step, head, (uint32_t)i));
}

// ---- first-layer q | k | v of the fast stack for every code ------------------------------------------------------------
// Passes >= 1 of the fast stack start from the embedding of a code (inference.py:123-149), so their first wqkv product is a
// function of that code alone: table[code] = bf16(wqkv . rmsnorm(fast_emb[code]) + bias), with the roundings of the decode
// kernel's staging (fp32 normalise, round, x weight, round) and an fp32 dot product per row.  Built once at finalize.
__global__ void __launch_bounds__(256) t0_build_kernel(const bf16 *emb, const bf16 *norm_w, const bf16 *W, const bf16 *bias, bf16 *table,
                                                       int K, int rows, float eps) {
  extern __shared__ __align__(16) float t0_xs[];      // K floats
  __shared__ float red[40];
  const int code = blockIdx.x, tid = threadIdx.x;
  float ss = 0.f;
  for (int e = tid; e < K; e += blockDim.x) { const float v = bf2f(emb[(size_t)code * K + e]); t0_xs[e] = v; ss = fmaf(v, v, ss); }
  ss = block_sum(ss, red);
  const float inv = rsqrtf(ss * (1.0f / (float)K) + eps);
  for (int e = tid; e < K; e += blockDim.x) t0_xs[e] = rbf(__fmul_rn(rbf(__fmul_rn(t0_xs[e], inv)), bf2f(norm_w[e])));
  __syncthreads();
  for (int r = tid; r < rows; r += blockDim.x) {
    const uint4 *wr = reinterpret_cast<const uint4 *>(W + (size_t)r * K);
    float acc = 0.f;
    for (int c = 0; c < (K >> 3); ++c) {
      const uint4 u = wr[c];
      const float *x = t0_xs + 8 * c;
      acc = fmaf(__uint_as_float(u.x << 16), x[0], acc); acc = fmaf(__uint_as_float(u.x & 0xffff0000u), x[1], acc);
      acc = fmaf(__uint_as_float(u.y << 16), x[2], acc); acc = fmaf(__uint_as_float(u.y & 0xffff0000u), x[3], acc);
      acc = fmaf(__uint_as_float(u.z << 16), x[4], acc); acc = fmaf(__uint_as_float(u.z & 0xffff0000u), x[5], acc);
      acc = fmaf(__uint_as_float(u.w << 16), x[6], acc); acc = fmaf(__uint_as_float(u.w & 0xffff0000u), x[7], acc);
    }
    if (bias) acc += bf2f(bias[r]);
    table[(size_t)code * rows + r] = f2bf(acc);
  }
}

}  // namespace da
