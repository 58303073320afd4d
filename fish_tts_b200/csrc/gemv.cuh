// gemv.cuh -- the weight-streaming bf16 GEMV family: every nn.Linear of the decode step at batch 1.
//
// Replaces (reference fish_tts/models/llama.py): wqkv :240, wo :283, w1/w3/w2 :190, the LM head
// :446-451 and fast_output :577-578, fused with what surrounds them:
//   prologue  PLAIN     x -> fp32 in shared memory
//             RMSNORM   the custom RMSNorm :172-177 (fp32 normalise, round to bf16, THEN * weight)
//             FASTATTN  the whole fast-layer attention :246-251, 285-309 (RoPE, KV write, bf16
//                       scores/softmax/PV over <= num_codebooks positions), recomputed per CTA
//   epilogue  STORE     bf16(acc + bias)
//             RESIDUAL  bf16(res + bf16(acc + bias))                                   :329-330
//             SWIGLU    rows come interleaved (w1_j, w3_j): bf16(bf16(silu(a)) * b)    :190
//             LOGITS    repetition penalty on the penalised ids, logits to L2-resident global
//                       memory, per-CTA (max, sum exp) partials; the fast heads finish the whole
//                       sampling step in the last CTA to arrive.
//
// Work split: rows are dealt to warps in adjacent PAIRS, pair p -> CTA p % grid, warp p / grid, so
// all SMs stream equal shares; lane l of segment s holds the 8 weights [s*256 + l*8, +8) as one 128-bit
// load (L1 bypassed, explicit L2 eviction policy).  CANONICAL ORDER of a row's dot product, shared by
// every kernel so that all paths agree bit for bit: per lane, each batch of DA_CH=4 segments is an fp32
// fma chain starting from zero; the batch partials are added in batch order; then a 5-step butterfly.
#pragma once
#include "common.cuh"
#include "sampler.cuh"

namespace da {

enum { PRO_PLAIN = 0, PRO_RMSNORM = 1, PRO_FASTATTN = 2 };
enum { EPI_STORE = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_LOGITS = 3 };

struct FastAttnArgs {
  const bf16 *qkv;        // [(nh + 2 nkv) * hd], output of this pass's wqkv
  bf16 *kc, *vc;          // [nkv][ncb][hd]
  const bf16 *rope;       // [ncb][hd/2][2] bf16 (llama.py:537-541)
  const bf16 *qn, *kn;    // optional qk-norm weights [hd]
  int nh, nkv, hd, ncb, pos;
  float eps, scale;       // scale = 1/sqrt(hd)
};

struct GemvArgs {
  const bf16 *W;          // [rows][K] row-major
  const bf16 *bias;       // [rows] or null
  int rows, K;
  int evict_last;         // L2 policy for W: 1 = keep (fast stack), 0 = stream
  const bf16 *x;          // prologue input [K]
  const bf16 *norm_w;     // RMSNORM weight [K]
  float eps;
  FastAttnArgs fa;
  const bf16 *res;        // RESIDUAL input [rows]
  bf16 *out;              // STORE/RESIDUAL: [rows]; SWIGLU: [rows/2]; LOGITS: logits [rows]
  // LOGITS
  bf16 *logits_raw;       // optional copy before the penalty
  float2 *partials;       // [grid] (max, sumexp)
  int head;               // 0 = slow head, k >= 1 = fast head of codebook k
  long long noise_off;    // offset of this head inside one step's noise block
  const bf16 *fast_emb;   // [codebook_size][fast_dim]
  bf16 *fast_x;           // next fast pass input [fast_dim]
  int fast_dim, last_head, codebook_size;
  int *seq; int seq_stride; int im_end_id; int n_rows_tok;  // finish_step
  int park_on_done;       // request slots: a finished request's position becomes -1, so the steps it still rides along in skip its KV cache
                          // (the slot may already be receiving the next request's prefill)
  DAState *st;
  Timeline tl;
};

// one warp applies (optional) nn.RMSNorm and RoPE to one head vector held as fp32 in shared memory
// llama.py:246-251 + 606-618.  nn.RMSNorm rounds once, after the weight multiply.
__device__ __forceinline__ void head_norm_rope(float *v, int hd, const bf16 *nw, float eps, const bf16 *rope_row, int lane) {
  if (nw) {
    float ss = 0.f;
    for (int d = lane; d < hd; d += 32) ss = fmaf(v[d], v[d], ss);
    ss = warp_sum(ss);
    float rstd = rsqrtf(ss * (1.0f / (float)hd) + eps);
    for (int d = lane; d < hd; d += 32) v[d] = rbf(__fmul_rn(__fmul_rn(v[d], rstd), bf2f(nw[d])));
    __syncwarp();
  }
  for (int i = lane; i < (hd >> 1); i += 32) {
    float x0 = v[2 * i], x1 = v[2 * i + 1];
    float c = bf2f(rope_row[2 * i]), s = bf2f(rope_row[2 * i + 1]);
    // separate mul / mul / add kernels in the reference: no fma contraction
    float o0 = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
    float o1 = __fadd_rn(__fmul_rn(x1, c), __fmul_rn(x0, s));
    v[2 * i] = rbf(o0); v[2 * i + 1] = rbf(o1);
  }
  __syncwarp();
}

// ---- prologues: fill xs (fp32, xs_index layout) ------------------------------------------------
// Activation vectors are moved as 16-byte chunks (8 bf16), every load of a phase issued before the first use, so a
// prologue costs ONE L2 round trip instead of K/256 dependent ones.  Chunk c -> thread c % blockDim, slot c / blockDim.
#define DA_XCH 2   // chunks per thread: K <= 8 * 256 * DA_XCH = 4096 with 256 threads
struct XRegs { uint4 v[DA_XCH]; };

__device__ __forceinline__ void load_chunks(XRegs &r, const bf16 *x, int K) {
#pragma unroll
  for (int i = 0; i < DA_XCH; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c * 8 < K) r.v[i] = *reinterpret_cast<const uint4 *>(x + (size_t)c * 8);
  }
}
// chunk c (elements 8c..8c+7) in the xs_index layout: two conflict-free float4 stores
__device__ __forceinline__ void store_chunk_xs(float *xs, int c, const float *f) {
  float4 *dst = reinterpret_cast<float4 *>(xs);
  const int sgm = c >> 5, lane = c & 31;
  dst[(sgm * 2 + 0) * 32 + lane] = make_float4(f[0], f[1], f[2], f[3]);
  dst[(sgm * 2 + 1) * 32 + lane] = make_float4(f[4], f[5], f[6], f[7]);
}

__device__ __forceinline__ void prologue_plain(const GemvArgs &a, float *xs, const XRegs &xr) {
#pragma unroll
  for (int i = 0; i < DA_XCH; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c * 8 < a.K) { float f[8]; unpack8(xr.v[i], f); store_chunk_xs(xs, c, f); }
  }
}

// wn: the norm weight chunks (a weight: fetched before the dependency wait); xr: the activation chunks
__device__ __forceinline__ void prologue_rmsnorm(const GemvArgs &a, float *xs, float *scratch, const XRegs &xr, const XRegs &wn) {
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < DA_XCH; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c * 8 < a.K) {
      float f[8]; unpack8(xr.v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
    }
  }
  ss = block_sum(ss, scratch);
  const float inv = rsqrtf(ss * (1.0f / (float)a.K) + a.eps);
#pragma unroll
  for (int i = 0; i < DA_XCH; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c * 8 < a.K) {
      float f[8], g[8]; unpack8(xr.v[i], f); unpack8(wn.v[i], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = rbf(__fmul_rn(rbf(__fmul_rn(f[j], inv)), g[j]));   // .type_as(x), then * weight
      store_chunk_xs(xs, c, f);
    }
  }
}

// fast-layer attention for ONE query position, every CTA recomputes it (<= 16 heads x <= 15 positions)
// work: [q | k_all | v_all | p] floats after xs
#define DA_FKV 4   // chunks of earlier K (and V) rows per thread: pos * nkv * hd <= 8 * 256 * DA_FKV
__device__ __forceinline__ void prologue_fastattn(const GemvArgs &a, float *xs, float *work) {
  const FastAttnArgs &f = a.fa;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int qd = f.nh * f.hd, kd = f.nkv * f.hd, P = f.pos + 1, G = f.nh / f.nkv;
  float *q = work, *ka = q + qd, *va = ka + f.ncb * kd, *pr = va + f.ncb * kd;   // pr: [nh][ncb]
  // one round trip: this position's q|k|v and every earlier K/V row (cache layout [nkv][ncb][hd], tiny, L2-resident)
  XRegs cur; load_chunks(cur, f.qkv, qd + 2 * kd);
  uint4 pk[DA_FKV], pv[DA_FKV];
  const int hd8 = f.hd >> 3, kd8 = kd >> 3, nprev = f.pos * kd8;
#pragma unroll
  for (int i = 0; i < DA_FKV; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < nprev) {
      const int j = c / kd8, r = c - j * kd8, g = r / hd8, d8 = r - g * hd8;
      const size_t src = ((size_t)g * f.ncb + j) * f.hd + (size_t)d8 * 8;
      pk[i] = *reinterpret_cast<const uint4 *>(f.kc + src);
      pv[i] = *reinterpret_cast<const uint4 *>(f.vc + src);
    }
  }
#pragma unroll
  for (int i = 0; i < DA_XCH; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c * 8 < qd + 2 * kd) {
      float t[8]; unpack8(cur.v[i], t);
      const int e = c * 8;
      float *dst = e < qd ? q + e : (e < qd + kd ? ka + f.pos * kd + (e - qd) : va + f.pos * kd + (e - qd - kd));
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = t[j];
    }
  }
#pragma unroll
  for (int i = 0; i < DA_FKV; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < nprev) {
      float t[8];
      unpack8(pk[i], t);
#pragma unroll
      for (int j = 0; j < 8; ++j) ka[c * 8 + j] = t[j];
      unpack8(pv[i], t);
#pragma unroll
      for (int j = 0; j < 8; ++j) va[c * 8 + j] = t[j];
    }
  }
  __syncthreads();
  const bf16 *rope_row = f.rope + (size_t)f.pos * f.hd;
  for (int h = w; h < f.nh + f.nkv; h += nw) {
    if (h < f.nh) head_norm_rope(q + h * f.hd, f.hd, f.qn, f.eps, rope_row, lane);
    else head_norm_rope(ka + f.pos * kd + (h - f.nh) * f.hd, f.hd, f.kn, f.eps, rope_row, lane);
  }
  __syncthreads();
  if (blockIdx.x == 0) {   // KVCache.update (llama.py:142-149) -- one writer
    for (int e = threadIdx.x; e < kd; e += blockDim.x) {
      int g = e / f.hd, d = e - g * f.hd;
      size_t dst = ((size_t)g * f.ncb + f.pos) * f.hd + d;
      f.kc[dst] = f2bf(ka[f.pos * kd + e]); f.vc[dst] = f2bf(va[f.pos * kd + e]);
    }
  }
  // scores: bf16(q @ k^T) then bf16(* scale)   (llama.py:304)
  for (int t = threadIdx.x; t < f.nh * P; t += blockDim.x) {
    int h = t / P, j = t - h * P, g = h / G;
    const float4 *qq = reinterpret_cast<const float4 *>(q + h * f.hd), *kk = reinterpret_cast<const float4 *>(ka + j * kd + g * f.hd);
    float acc = 0.f;
    for (int d = 0; d < (f.hd >> 2); ++d) {
      float4 x = qq[d], y = kk[d];
      acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
    }
    pr[h * f.ncb + j] = rbf(__fmul_rn(rbf(acc), f.scale));
  }
  __syncthreads();
  // softmax over j <= pos in fp32, rounded to bf16 (llama.py:305-306; masked columns are exp(-inf)=0).
  // one thread per (h, j); every thread of a row walks it in the same order, so the row sum is identical
  float pval = 0.f;
  const int t = threadIdx.x;
  if (t < f.nh * P) {
    const int h = t / P, j = t - h * P;
    float m = -INFINITY;
    for (int jj = 0; jj < P; ++jj) m = fmaxf(m, pr[h * f.ncb + jj]);
    float sum = 0.f;
    for (int jj = 0; jj < P; ++jj) sum += expf(pr[h * f.ncb + jj] - m);
    pval = rbf(expf(pr[h * f.ncb + j] - m) / sum);
  }
  __syncthreads();
  if (t < f.nh * P) { const int h = t / P, j = t - h * P; pr[h * f.ncb + j] = pval; }
  __syncthreads();
  // y = bf16(p @ v)  (llama.py:309), laid out [h*hd + d] = the wo input; one 8-element chunk per thread
  for (int c = threadIdx.x; c * 8 < qd; c += blockDim.x) {
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
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = rbf(acc[j]);
    store_chunk_xs(xs, c, acc);
  }
}

// ---- the dot core: two adjacent rows per warp, DA_CH segments (of 256 elements) per batch ------------
__device__ __forceinline__ void fma8(const uint4 &wv, const float4 &x0, const float4 &x1, float &acc) {
  float f[8]; unpack8(wv, f);
  acc = fmaf(f[0], x0.x, acc); acc = fmaf(f[1], x0.y, acc); acc = fmaf(f[2], x0.z, acc); acc = fmaf(f[3], x0.w, acc);
  acc = fmaf(f[4], x1.x, acc); acc = fmaf(f[5], x1.y, acc); acc = fmaf(f[6], x1.z, acc); acc = fmaf(f[7], x1.w, acc);
}

#define DA_CH 4   // 8 x 128-bit loads in flight per lane per batch; two batches are kept in flight (ping-pong)

struct PairRegs { uint4 u0[DA_CH], u1[DA_CH]; };

// batch b of row pair p: segments [b*DA_CH, +DA_CH) of rows 2p and 2p+1
__device__ __forceinline__ void load_batch(PairRegs &r, const bf16 *__restrict__ W, int K, int rows, int p, int b, int nseg,
                                           int lane, uint64_t pol) {
  const int r0 = 2 * p;
  const bool has1 = r0 + 1 < rows;
  const uint4 *p0 = reinterpret_cast<const uint4 *>(W + (size_t)r0 * K) + lane;
  const uint4 *p1 = reinterpret_cast<const uint4 *>(W + (size_t)(has1 ? r0 + 1 : r0) * K) + lane;
#pragma unroll
  for (int c = 0; c < DA_CH; ++c) {
    const int sgm = b * DA_CH + c;
    if (sgm < nseg) { r.u0[c] = ldg_w(p0 + sgm * 32, pol); r.u1[c] = ldg_w(p1 + sgm * 32, pol); }
  }
}
__device__ __forceinline__ void fma_batch(const PairRegs &r, const float *xs, int b, int nseg, int lane, float &a0, float &a1) {
  const float4 *xv = reinterpret_cast<const float4 *>(xs) + lane;
#pragma unroll
  for (int c = 0; c < DA_CH; ++c) {
    const int sgm = b * DA_CH + c;
    if (sgm < nseg) {
      float4 x0 = xv[(sgm * 2 + 0) * 32], x1 = xv[(sgm * 2 + 1) * 32];
      fma8(r.u0[c], x0, x1, a0);
      fma8(r.u1[c], x0, x1, a1);
    }
  }
}

// programmatic dependent launch (guide: Guideline 9): the next kernel of the step graph is launched while this one
// runs; it may only touch weights until griddepcontrol.wait returns (= every earlier kernel has completed).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- end of a decode step (loop mode): what decode_n_tokens does between calls -------------------
// inference.py:186-211: record the column, advance input_pos, rebuild the 16-wide window, EOS test.
__device__ __forceinline__ void finish_step(const GemvArgs &a) {
  DAState *st = a.st;
  if (!st->loop_mode) return;
  const int R = a.n_rows_tok;
  int pos = st->pos + 1, n_gen = st->n_gen + 1;
  for (int r = 0; r < R; ++r) { a.seq[(size_t)r * a.seq_stride + pos] = st->tok_out[r]; st->tok_in[r] = st->tok_out[r]; }
  st->pos = pos; st->n_gen = n_gen; st->step_ctr += 1; st->use_penalty = 1;
  if (st->noise) st->noise += st->noise_stride;
  // the reference tests for <|im_end|> only inside decode_n_tokens (inference.py:210): the column produced by the
  // prefill call is never checked, so an EOS there is followed by one more step
  if ((n_gen > 1 && st->tok_out[0] == a.im_end_id) || n_gen >= st->max_gen) { st->done = 1; if (a.park_on_done) st->pos = -1; }
  // previous_tokens[:, j] = generated column j+1 (the prefill-produced column 0 is never recorded)
  int i = n_gen - 1, T = st->prompt_len;
  for (int c = 0; c < DA_WIN; ++c) {
    int j = i < DA_WIN ? c : i - DA_WIN + c;
    for (int r = 0; r < R; ++r)
      st->win[r * DA_WIN + c] = (j < i) ? a.seq[(size_t)r * a.seq_stride + T + 1 + j] : 0;
  }
}

// the fast heads (<= 1024 logits): one CTA samples and prepares the next pass
#define DA_FAST_IPT 4   // 256 threads x 4 items
__device__ void fast_head_sample(const GemvArgs &a, float *smem) {
  DAState *st = a.st;
  const int V = a.rows;
  unsigned long long *scr = reinterpret_cast<unsigned long long *>(smem);   // 192 u64
  float *scrf = reinterpret_cast<float *>(scr + 192);
  const volatile uint16_t *lg = reinterpret_cast<const volatile uint16_t *>(a.out);
  uint32_t key[DA_FAST_IPT], idx[DA_FAST_IPT], valid = 0;
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < DA_FAST_IPT; ++i) {
    const int e = threadIdx.x * DA_FAST_IPT + i;
    idx[i] = (uint32_t)e; key[i] = 0;
    if (e < V) { const uint16_t b = lg[e]; key[i] = bf16_key(b); valid |= 1u << i; mx = fmaxf(mx, bits2f(b)); }
  }
  SampleParams sp;
  sp.m = block_max(mx, scrf);
  {   // sum of exp terms as 2^-40 fixed point: order-free, so every path computes the same S
    Red es = {0ull, 0, -1}; int par = 0;
#pragma unroll
    for (int i = 0; i < DA_FAST_IPT; ++i) if ((valid >> i) & 1u) es.s += (unsigned long long)(expf(bits2f(key_bf16(key[i])) - sp.m) * DA_FIX2_SCALE);
    sp.S = __ull2float_rn(block_reduce(es, scr, par).s) * (1.0f / DA_FIX2_SCALE);
    __syncthreads();
  }
  sp.T_bf = eff_temperature(st);
  sp.c_max = cmax_from_top_p(st->top_p);
  uint32_t tok = sample_items<DA_FAST_IPT>(key, idx, valid, (uint32_t)V, true, sp, noise_src(st), (uint32_t)a.head, a.noise_off, &st->nucleus[a.head], scr);
  if (tok >= (uint32_t)a.codebook_size) { tok = a.codebook_size - 1; if (threadIdx.x == 0) st->err = 3; }
  for (int d = threadIdx.x; d < a.fast_dim; d += blockDim.x) a.fast_x[d] = a.fast_emb[(size_t)tok * a.fast_dim + d];
  __syncthreads();
  if (threadIdx.x == 0) {
    st->tok_out[a.head + 1] = (int)tok;
    if (a.last_head) finish_step(a);
  }
}

// the last CTA of the fast-head GEMV to arrive runs it
__device__ void fast_head_tail(const GemvArgs &a, float *smem) {
  DAState *st = a.st;
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->fast_ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) st->fast_ticket = 0;
  fast_head_sample(a, smem);
}

// ---- the kernel ------------------------------------------------------------------------------
// dynamic shared memory: xs[K] floats | 80 floats scratch | prologue / sampler workspace
// A warp walks "units" = (row pair, batch of DA_CH segments); unit u+1 is in flight while unit u is consumed, and the
// first unit is fetched BEFORE the dependency wait, so the weight stream of kernel N+1 overlaps the tail of kernel N.
#define DA_GEMV_THREADS 256
template <int PRO, int EPI>
__global__ void __launch_bounds__(DA_GEMV_THREADS, 2) gemv_kernel(const GemvArgs a) {
  extern __shared__ __align__(16) float smem[];
  tl_stamp(a.tl, 0);
  pdl_launch_dependents();
  float *xs = smem;
  float *scratch = smem + a.K;
  float *work = scratch + 80;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint64_t pol = a.evict_last ? policy_evict_last() : policy_evict_first();
  const int nseg = a.K >> 8;
  const int nb = (nseg + DA_CH - 1) / DA_CH;
  const int npairs = (a.rows + 1) >> 1;
  const int first = w * gridDim.x + blockIdx.x, stride = nw * gridDim.x;
  const int my_pairs = first < npairs ? (npairs - first + stride - 1) / stride : 0;
  const int nu = my_pairs * nb;

  PairRegs ra, rb;
  if (nu > 0) load_batch(ra, a.W, a.K, a.rows, first, 0, nseg, lane, pol);
  XRegs wn, xr;
  if (PRO == PRO_RMSNORM) load_chunks(wn, a.norm_w, a.K);

  pdl_wait();
  tl_stamp(a.tl, 1);
  // everything this kernel needs from its predecessors is requested in one go: the done flag, the activation
  // vector and (RESIDUAL) the residual values of this warp's first row pair
  const int done = *reinterpret_cast<const volatile int *>(&a.st->done);
  if (PRO != PRO_FASTATTN) load_chunks(xr, a.x, a.K);
  float res0 = 0.f, res1 = 0.f;
  if (EPI == EPI_RESIDUAL && nu > 0 && lane == 0) {
    res0 = bf2f(a.res[2 * first]);
    if (2 * first + 1 < a.rows) res1 = bf2f(a.res[2 * first + 1]);
  }
  if (done) return;

  if (PRO == PRO_PLAIN) prologue_plain(a, xs, xr);
  else if (PRO == PRO_RMSNORM) prologue_rmsnorm(a, xs, scratch, xr, wn);
  else prologue_fastattn(a, xs, work);
  __syncthreads();
  tl_stamp(a.tl, 2);

  // LOGITS: penalised ids and the online (max, sumexp) of this warp's rows
  int pen_id = -1; float rp_bf = 1.f; float wm = -INFINITY, wl = 0.f;
  if (EPI == EPI_LOGITS) {
    const DAState *st = a.st;
    if (st->use_penalty) {
      rp_bf = eff_rep_penalty(st);
      if (a.head == 0) { if (lane < a.n_rows_tok) pen_id = st->win[lane * DA_WIN]; }       // previous_tokens[:, 0]  (inference.py:109-111)
      else if (lane < DA_WIN) pen_id = st->win[(a.head + 1) * DA_WIN + lane];               // previous_tokens[k+1]   (inference.py:141-145)
    }
  }

  float a0 = 0.f, a1 = 0.f;
  int pi = 0, b = 0;   // unit u = (pair index pi, batch b)
  auto step = [&](PairRegs &cur, PairRegs &nxt, int u) {
    // prefetch unit u+1 into the other register set
    int pin = pi, bn = b + 1;
    if (bn == nb) { bn = 0; ++pin; }
    if (u + 1 < nu) load_batch(nxt, a.W, a.K, a.rows, first + pin * stride, bn, nseg, lane, pol);
    {   // canonical order: each batch of DA_CH segments is a chain from zero; batch partials fold in batch order
      float p0 = 0.f, p1 = 0.f;
      fma_batch(cur, xs, b, nseg, lane, p0, p1);
      a0 += p0; a1 += p1;
    }
    if (b == nb - 1) {
      const int p = first + pi * stride;
      const int r0 = 2 * p, r1 = r0 + 1;
      const bool has1 = r1 < a.rows;
      float d0 = warp_sum(a0), d1 = warp_sum(a1);
      a0 = 0.f; a1 = 0.f;
      if (a.bias) { d0 += bf2f(a.bias[r0]); if (has1) d1 += bf2f(a.bias[r1]); }
      if (EPI == EPI_STORE) {
        if (lane == 0) { a.out[r0] = f2bf(d0); if (has1) a.out[r1] = f2bf(d1); }
      } else if (EPI == EPI_RESIDUAL) {
        if (lane == 0) {
          if (pi > 0) { res0 = bf2f(a.res[r0]); if (has1) res1 = bf2f(a.res[r1]); }
          a.out[r0] = f2bf(res0 + rbf(d0));
          if (has1) a.out[r1] = f2bf(res1 + rbf(d1));
        }
      } else if (EPI == EPI_SWIGLU) {
        if (lane == 0) {
          float g = rbf(d0), up = rbf(d1);
          float sg = rbf(g / (1.0f + expf(-g)));       // F.silu in fp32, rounded
          a.out[p] = f2bf(__fmul_rn(sg, up));
        }
      } else {
        float z0 = rbf(d0), z1 = rbf(d1);
        if (a.logits_raw && lane == 0) { a.logits_raw[r0] = f2bf(z0); if (has1) a.logits_raw[r1] = f2bf(z1); }
        unsigned hit0 = __ballot_sync(0xffffffffu, pen_id == r0), hit1 = __ballot_sync(0xffffffffu, pen_id == r1);
        if (hit0) z0 = penalise(z0, rp_bf);
        if (hit1) z1 = penalise(z1, rp_bf);
        if (lane == 0) {
          a.out[r0] = f2bf(z0);
          float mn = fmaxf(wm, z0); wl = wl * expf(wm - mn) + expf(z0 - mn); wm = mn;
          if (has1) {
            a.out[r1] = f2bf(z1);
            mn = fmaxf(wm, z1); wl = wl * expf(wm - mn) + expf(z1 - mn); wm = mn;
          }
        }
      }
    }
    pi = pin; b = bn;
  };
  for (int u = 0; u < nu; u += 2) {   // ping-pong with static register sets
    step(ra, rb, u);
    if (u + 1 < nu) step(rb, ra, u + 1);
  }

  tl_stamp(a.tl, 3);
  if (EPI == EPI_LOGITS) {
    if (a.head == 0) {
      // per-CTA partial of the softmax statistics; combined by the select kernel
      __syncthreads();
      if (lane == 0) { scratch[w] = wm; scratch[40 + w] = wl; }
      __syncthreads();
      if (w == 0) {
        float mi = lane < nw ? scratch[lane] : -INFINITY, li = lane < nw ? scratch[40 + lane] : 0.f;
        float m = warp_max(mi);
        float l = warp_sum(li > 0.f ? li * expf(mi - m) : 0.f);
        if (lane == 0) a.partials[blockIdx.x] = make_float2(m, l);
      }
    } else {
      fast_head_tail(a, work);
    }
  }
}

}  // namespace da
