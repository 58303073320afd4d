// persistent.cuh -- a table-driven PERSISTENT phase machine: many dependent GEMV phases in ONE cooperative kernel.
//
// Used for the fast-AR codebook loop (inference.py:115-149: 10 passes x 4 layers x {wqkv, attention+wo, w1/w3, w2} + 9 heads
// with their samplers = 169 phases per token on s1-mini).  As separate kernels every phase pays a launch boundary
// (3-4 us measured, profiles/r01_timeline_per_phase_kernels.txt); here one CTA per SM stays resident (cooperative
// launch guarantees co-residency) and phases hand over through memory alone:
//   * every activation element travels as one 32-bit UNIT = bf16 value (high half) | 16-bit phase tag (low half),
//     written with a single 4-byte store; a consumer polls until the tag matches the producing phase.  Data and
//     "ready" arrive atomically: no fence, no atomic counter, no grid barrier between phases (the idea of NCCL's LL
//     protocol, on-chip through L2).  Tags advance with a per-engine phase counter, so consecutive uses of a buffer
//     never share a tag and nothing needs clearing.
//   * polling etiquette: unthrottled all-thread polling saturates L2 and starves the weight stream (measured 2x
//     slowdown), so ONE thread per CTA watches a sentinel unit with back-off before the CTA reads the vector.
//   * the schedule is a table of PhaseDesc in shared memory; each warp keeps the first 128-bit batch of its NEXT
//     work unit in flight, possibly several phases ahead, so the weight stream never waits for a hand-over.
//   * narrow phases (few rows, long K: wo, w2) split K across warps by batch; lane partials meet in shared memory and
//     fold in batch order -- the same canonical order as gemv.cuh, so both paths agree bit for bit.
//   * the fast KV cache of the current token lives in shared memory (every CTA recomputes the <=16-head attention
//     of one position), so no cross-CTA ordering is needed for it.
// All spins are bounded: a lost hand-over raises the device fault flag instead of hanging the GPU.
#pragma once
#include "common.cuh"
#include "gemv.cuh"
#include "sampler.cuh"

namespace da {

#define DA_P_THREADS 512
#define DA_SPIN_LIMIT (1 << 22)
#define DA_MAX_FAST_LAYERS 8
#define DA_PART_UNITS 64     // K-split partial slots per CTA per phase (pairs_per_cta * batches)

enum { PP_PLAIN = 0, PP_RMSNORM = 1, PP_FASTATTN = 2 };
enum { PE_STORE = 0, PE_RESIDUAL = 1, PE_SWIGLU = 2, PE_FASTLOGITS = 3 };

struct PhaseDesc {
  const bf16 *W, *bias, *norm_w;
  const void *in;        // uint32_t units or plain bf16, see in_units
  const void *res;       // residual vector (RESIDUAL), units or plain
  uint32_t *out;         // output units
  int rows, K;
  int in_ph;             // index of the phase that wrote `in` (its tag)
  short pro, epi, in_units, res_units, layer, pos, evict_last, pad;
};

struct PersistArgs {
  const PhaseDesc *table; int n_phases;
  // fast attention
  const bf16 *rope; const bf16 *qn[DA_MAX_FAST_LAYERS]; const bf16 *kn[DA_MAX_FAST_LAYERS];
  int n_layer, nh, nkv, hd, ncb;
  float eps, scale;
  // fast heads
  const bf16 *fast_emb; int dim, fv, codebook_size;
  uint32_t *u_fin; bf16 *flogits_raw, *flogits; long long noise_off0;
  // end of step
  int *seq; int seq_stride, im_end_id, n_rows_tok;
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
__device__ __forceinline__ bool wait_sentinel(const uint32_t *p, uint32_t tag) {
  int it = 0;
  while ((ld_poll1(p) & 0xFFFFu) != tag) { if (++it >= DA_SPIN_LIMIT) return false; __nanosleep(60); }
  return true;
}
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// two-barrier block sum (all threads get the result); scratch >= 32 floats, distinct from the previous call's by parity
__device__ __forceinline__ float block_sum2(float v, float *scratch, int nw, int lane, int w) {
  v = warp_sum(v);
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += scratch[i];   // fixed order
  return t;
}

// ---- sampling of one fast head by warps 0..3 of CTA 0 (128 threads x 8 items, named barrier 1) ------------------------------
struct Red4 { unsigned long long s; int c; int m; };
__device__ __forceinline__ Red4 group_reduce(Red4 v, unsigned long long *scr, int &parity, int lane, int w) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.s += __shfl_xor_sync(0xffffffffu, v.s, o);
    v.c += __shfl_xor_sync(0xffffffffu, v.c, o);
    v.m = max(v.m, __shfl_xor_sync(0xffffffffu, v.m, o));
  }
  unsigned long long *b = scr + parity * 12;
  parity ^= 1;
  if (lane == 0) { b[w] = v.s; b[4 + w] = (unsigned long long)(long long)v.c; b[8 + w] = (unsigned long long)(long long)v.m; }
  named_bar(1, 128);
  Red4 r = {0ull, 0, INT_MIN};
#pragma unroll
  for (int i = 0; i < 4; ++i) { r.s += b[i]; r.c += (int)(long long)b[4 + i]; r.m = max(r.m, (int)(long long)b[8 + i]); }
  return r;
}

// same definition as sample_items (sampler.cuh), for a 128-thread group; all items present (whole fast vocabulary)
#define DA_G_IPT 8
__device__ uint32_t sample_group(const uint32_t (&key)[DA_G_IPT], const uint32_t (&idx)[DA_G_IPT], uint32_t valid_mask, uint32_t idx_limit,
                                 const SampleParams &sp, const DAState *st, uint32_t head, long long head_off, int *nucleus_out,
                                 unsigned long long *scr, int lane, int w) {
  int parity = 0;
  unsigned long long wt[DA_G_IPT];
  Red4 t = {0ull, 0, -1};
  int kmin = 65536;
#pragma unroll
  for (int i = 0; i < DA_G_IPT; ++i) {
    const bool v = (valid_mask >> i) & 1u;
    wt[i] = v ? pweight(bits2f(key_bf16(key[i])), sp.m, sp.S) : 0ull;
    t.s += wt[i]; t.c += v; if (v) { t.m = max(t.m, (int)key[i]); kmin = min(kmin, (int)key[i]); }
  }
  const Red4 tot = group_reduce(t, scr, parity, lane, w);
  Red4 mn = {0ull, 0, -kmin};
  const int key_lo = -group_reduce(mn, scr, parity, lane, w).m;      // smallest key present
  const int top_key = tot.m;
  // (1) lowest key kappa with G(kappa) = sum_{key >= kappa} w <= c_max; G(top_key + 1) = 0, G(key_lo) = total
  uint32_t lo = (uint32_t)key_lo, hi = (uint32_t)top_key + 1u;
  if (tot.s <= sp.c_max) hi = 0;
  else while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    Red4 g = {0ull, 0, -1};
#pragma unroll
    for (int i = 0; i < DA_G_IPT; ++i) if (((valid_mask >> i) & 1u) && key[i] >= mid) g.s += wt[i];
    if (group_reduce(g, scr, parity, lane, w).s <= sp.c_max) hi = mid; else lo = mid;
  }
  const uint32_t kappa = hi;
  Red4 g = {0ull, 0, -1};
#pragma unroll
  for (int i = 0; i < DA_G_IPT; ++i) if ((valid_mask >> i) & 1u) {
    if (key[i] >= kappa) { g.s += wt[i]; g.c += 1; } else g.m = max(g.m, (int)key[i]);
  }
  g = group_reduce(g, scr, parity, lane, w);
  const int n_full = g.c, tau = g.m;
  int c_part = 0; long long i_cut = -1;
  if (tau >= 0) {
    Red4 q = {0ull, 0, -1};
#pragma unroll
    for (int i = 0; i < DA_G_IPT; ++i) if (((valid_mask >> i) & 1u) && (int)key[i] == tau) { q.c += 1; q.m = max(q.m, (int)idx[i]); }
    q = group_reduce(q, scr, parity, lane, w);
    const int n_tau = q.c;
    const unsigned long long wtau = pweight(bits2f(key_bf16((uint32_t)tau)), sp.m, sp.S);
    const unsigned long long room = sp.c_max >= g.s ? sp.c_max - g.s : 0ull;
    const unsigned long long c = wtau ? room / wtau : (unsigned long long)n_tau;
    c_part = (int)(c < (unsigned long long)n_tau ? c : (unsigned long long)n_tau);
    if (n_full == 0 && c_part < 1) c_part = 1;
    if (c_part == n_tau) i_cut = (long long)idx_limit;
    else if (c_part > 0) {
      long long l = -1, h = (long long)q.m;      // largest index in the group bounds the search
      while (h - l > 1) {
        const long long mid = (l + h) >> 1;
        Red4 z = {0ull, 0, -1};
#pragma unroll
        for (int i = 0; i < DA_G_IPT; ++i) if (((valid_mask >> i) & 1u) && (int)key[i] == tau && (long long)idx[i] <= mid) z.c += 1;
        if (group_reduce(z, scr, parity, lane, w).c >= c_part) h = mid; else l = mid;
      }
      i_cut = h;
    }
  }
  if (threadIdx.x == 0 && nucleus_out) *nucleus_out = n_full + c_part;
  const float mz = rbf(bits2f(key_bf16((uint32_t)top_key)) / sp.T_bf);
  float e2[DA_G_IPT]; uint32_t keep = 0;
  Red4 s2 = {0ull, 0, -1};
#pragma unroll
  for (int i = 0; i < DA_G_IPT; ++i) {
    const bool k = ((valid_mask >> i) & 1u) && (key[i] >= kappa || ((int)key[i] == tau && (long long)idx[i] <= i_cut));
    e2[i] = 0.f;
    if (k) { keep |= 1u << i; e2[i] = expf(rbf(bits2f(key_bf16(key[i])) / sp.T_bf) - mz); s2.s += (unsigned long long)(e2[i] * DA_FIX2_SCALE); }
  }
  const float S2 = __ull2float_rn(group_reduce(s2, scr, parity, lane, w).s) * (1.0f / DA_FIX2_SCALE);
  ArgBest best = {0.f, 0u};
#pragma unroll
  for (int i = 0; i < DA_G_IPT; ++i) if ((keep >> i) & 1u) {
    const float p2 = rbf(e2[i] / S2);
    ArgBest cnd = {rbf(p2 / noise_at(st, head, head_off, idx[i])), idx[i]};
    best = better(best, cnd);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgBest tt; tt.r = __shfl_xor_sync(0xffffffffu, best.r, o); tt.idx = __shfl_xor_sync(0xffffffffu, best.idx, o);
    best = better(best, tt);
  }
  float *fs = reinterpret_cast<float *>(scr + 24);
  uint32_t *is = reinterpret_cast<uint32_t *>(fs + 4);
  if (lane == 0) { fs[w] = best.r; is[w] = best.idx; }
  named_bar(1, 128);
  ArgBest b = {fs[0], is[0]};
#pragma unroll
  for (int i = 1; i < 4; ++i) { ArgBest tt = {fs[i], is[i]}; b = better(b, tt); }
  return b.idx;
}

// shared memory: table | xs[4096] f32 | scratch 2x32 f32 | q[nh*hd] f32 | kcur,vcur [nkv*hd] f32 | pr [nh*ncb] f32 |
//                partials [DA_PART_UNITS][64] f32 | sampler scratch 32 u64 | kv store bf16 [n_layer][ncb][2][nkv*hd]
static inline size_t persist_smem_bytes(int n_phases, int n_layer, int nh, int nkv, int hd, int ncb) {
  size_t f = (size_t)n_phases * sizeof(PhaseDesc);
  f = (f + 15) & ~(size_t)15;
  f += (size_t)(4096 + 64 + nh * hd + 2 * nkv * hd + ((nh * ncb + 3) & ~3) + DA_PART_UNITS * 64) * sizeof(float) + 32 * 8;
  f = (f + 15) & ~(size_t)15;
  return f + (size_t)n_layer * ncb * 2 * nkv * hd * sizeof(bf16) + 64;
}

__global__ void __launch_bounds__(DA_P_THREADS, 1) persistent_kernel(const PersistArgs a) {
  extern __shared__ __align__(16) unsigned char smem_p[];
  DAState *st = a.st;
  tl_stamp(a.tl, 0);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = DA_P_THREADS / 32;
  const int qd = a.nh * a.hd, kd = a.nkv * a.hd, G = a.nh / a.nkv;
  // ---- carve shared memory, copy the phase table -------------------------------------------------------------------
  PhaseDesc *tab = reinterpret_cast<PhaseDesc *>(smem_p);
  size_t off = ((size_t)a.n_phases * sizeof(PhaseDesc) + 15) & ~(size_t)15;
  float *xs = reinterpret_cast<float *>(smem_p + off);
  float *scratch = xs + 4096, *q = scratch + 64, *kcur = q + qd, *vcur = kcur + kd, *pr = vcur + kd;
  float *part = pr + ((a.nh * a.ncb + 3) & ~3);
  unsigned long long *scr = reinterpret_cast<unsigned long long *>(part + DA_PART_UNITS * 64);
  bf16 *kvs = reinterpret_cast<bf16 *>(scr + 32);
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(a.table);
    uint4 *dst = reinterpret_cast<uint4 *>(tab);
    const int n16 = (int)((size_t)a.n_phases * sizeof(PhaseDesc) / 16);
    for (int i = threadIdx.x; i < n16; i += DA_P_THREADS) dst[i] = src[i];
  }
  const int done = *reinterpret_cast<const volatile int *>(&st->done);
  const uint32_t tag_base = *reinterpret_cast<const volatile unsigned int *>(&st->phase_ctr);
  __syncthreads();
  if (done) return;
  auto tag_of = [&](int ph) { return (uint32_t)((tag_base + (uint32_t)ph) % 65535u) + 1u; };
  const int nph = a.n_phases, grid = gridDim.x, bid = blockIdx.x;
  const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
  const float rp_eff = eff_rep_penalty(st);
  const int use_pen = st->use_penalty;
  bool ok = true;

  // work units of this CTA in phase ph: pairs j*grid + bid, each split into nb batches; unit u = j*nb + b -> warp u % nw
  auto units_of = [&](int ph, int &nb_out) {
    const PhaseDesc &d = tab[ph];
    const int npairs = (d.rows + 1) >> 1;
    const int npc = bid < npairs ? (npairs - bid + grid - 1) / grid : 0;
    nb_out = ((d.K >> 8) + DA_CH - 1) / DA_CH;
    return npc * nb_out;
  };
  auto issue = [&](PairRegs &r, int ph, int u) {
    const PhaseDesc &d = tab[ph];
    const int nseg = d.K >> 8, nb = (nseg + DA_CH - 1) / DA_CH;
    const int j = u / nb, b = u - j * nb;
    load_batch(r, d.W, d.K, d.rows, j * grid + bid, b, nseg, lane, d.evict_last ? pol_keep : pol_stream);
  };
  // cursor of the unit held in `ra`
  PairRegs ra;
  int c_ph = 0, c_u = w; bool c_valid = false;
  auto seek = [&](int ph, int u) {      // first unit at or after (ph, u) that belongs to this warp
    for (; ph < nph; ++ph, u = w) { int nb; if (u < units_of(ph, nb)) { c_ph = ph; c_u = u; return true; } }
    return false;
  };
  c_valid = seek(0, w);
  if (c_valid) issue(ra, c_ph, c_u);
  tl_stamp(a.tl, 1);

  int sparity = 0;
  for (int ph = 0; ph < nph; ++ph) {
    const PhaseDesc d = tab[ph];
    const uint32_t tag = tag_of(ph), in_tag = tag_of(d.in_ph);
    const int nseg = d.K >> 8, nb = (nseg + DA_CH - 1) / DA_CH;
    const int npairs = (d.rows + 1) >> 1;
    const int npc = bid < npairs ? (npairs - bid + grid - 1) / grid : 0;
    const bool split = nb > 1;
    Timeline tp = a.tl; tp.slot = 192 + ph;
    if (192 + ph < 512) tl_stamp(tp, 0);
    float *sc = scratch + sparity * 32; sparity ^= 1;

    // ---- (A) stage the input vector into xs ---------------------------------------------------------------------------
    if (d.pro == PP_FASTATTN) {
      // fast-layer attention for position d.pos (llama.py:246-251, 285-309), recomputed by every CTA
      const int pos = d.pos, P = pos + 1;
      const uint32_t *in_u = reinterpret_cast<const uint32_t *>(d.in);
      bf16 *kv_l = kvs + (size_t)d.layer * a.ncb * 2 * kd;
      if (threadIdx.x == 0) ok = wait_sentinel(in_u + qd + 2 * kd - 1, in_tag) && ok;
      __syncthreads();
      {
        const int c = threadIdx.x;
        if (c * 8 < qd + 2 * kd) {
          float t[8];
          ok = poll_chunk(in_u + c * 8, in_tag, t) && ok;
          const int e = c * 8;
          float *dst = e < qd ? q + e : (e < qd + kd ? kcur + (e - qd) : vcur + (e - qd - kd));
          *reinterpret_cast<float4 *>(dst) = make_float4(t[0], t[1], t[2], t[3]);
          *reinterpret_cast<float4 *>(dst + 4) = make_float4(t[4], t[5], t[6], t[7]);
        }
      }
      __syncthreads();
      const bf16 *rope_row = a.rope + (size_t)pos * a.hd;
      for (int h = w; h < a.nh + a.nkv; h += nw) {
        if (h < a.nh) head_norm_rope(q + h * a.hd, a.hd, a.qn[d.layer], a.eps, rope_row, lane);
        else head_norm_rope(kcur + (h - a.nh) * a.hd, a.hd, a.kn[d.layer], a.eps, rope_row, lane);
      }
      __syncthreads();
      // this token's fast KV row -> shared-memory cache; scores for every (h, j <= pos): bf16(q @ k^T), then bf16(* scale)
      for (int e = threadIdx.x; e < kd; e += DA_P_THREADS) {
        kv_l[((size_t)pos * 2 + 0) * kd + e] = f2bf(kcur[e]);
        kv_l[((size_t)pos * 2 + 1) * kd + e] = f2bf(vcur[e]);
      }
      for (int t = threadIdx.x; t < a.nh * P; t += DA_P_THREADS) {
        const int h = t / P, j = t - h * P, g = h / G;
        const float *qq = q + h * a.hd;
        float acc = 0.f;
        if (j == pos) {
          const float *kk = kcur + g * a.hd;
          for (int dd = 0; dd < a.hd; ++dd) acc = fmaf(qq[dd], kk[dd], acc);
        } else {
          const uint4 *kk = reinterpret_cast<const uint4 *>(kv_l + ((size_t)j * 2 + 0) * kd + g * a.hd);
          for (int d8 = 0; d8 < (a.hd >> 3); ++d8) {
            float kf[8]; unpack8(kk[d8], kf);
#pragma unroll
            for (int x = 0; x < 8; ++x) acc = fmaf(qq[d8 * 8 + x], kf[x], acc);
          }
        }
        pr[h * a.ncb + j] = rbf(__fmul_rn(rbf(acc), a.scale));
      }
      __syncthreads();
      // softmax (fp32, rounded to bf16) fused with y = bf16(p @ v): one 8-wide output chunk per thread
      for (int c = threadIdx.x; c * 8 < qd; c += DA_P_THREADS) {
        const int e = c * 8, h = e / a.hd, dd = e - h * a.hd, g = h / G;
        float m = -INFINITY;
        for (int jj = 0; jj < P; ++jj) m = fmaxf(m, pr[h * a.ncb + jj]);
        float sum = 0.f;
        for (int jj = 0; jj < P; ++jj) sum += expf(pr[h * a.ncb + jj] - m);
        float acc[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) acc[x] = 0.f;
        for (int jj = 0; jj < P; ++jj) {
          const float pj = rbf(expf(pr[h * a.ncb + jj] - m) / sum);
          float vf[8];
          if (jj == pos) {
#pragma unroll
            for (int x = 0; x < 8; ++x) vf[x] = vcur[g * a.hd + dd + x];
          } else unpack8(*reinterpret_cast<const uint4 *>(kv_l + ((size_t)jj * 2 + 1) * kd + g * a.hd + dd), vf);
#pragma unroll
          for (int x = 0; x < 8; ++x) acc[x] = fmaf(pj, vf[x], acc[x]);
        }
#pragma unroll
        for (int x = 0; x < 8; ++x) acc[x] = rbf(acc[x]);
        store_chunk_xs(xs, c, acc);
      }
    } else {
      const int c = threadIdx.x;
      const bool mine = c * 8 < d.K;
      float v[8], g[8];
      if (mine && d.pro == PP_RMSNORM) unpack8(*reinterpret_cast<const uint4 *>(d.norm_w + (size_t)c * 8), g);
      if (d.in_units) {
        const uint32_t *in_u = reinterpret_cast<const uint32_t *>(d.in);
        if (threadIdx.x == 0) ok = wait_sentinel(in_u + d.K - 1, in_tag) && ok;
        __syncthreads();
        if (mine) ok = poll_chunk(in_u + c * 8, in_tag, v) && ok;
      } else if (mine) unpack8(*reinterpret_cast<const uint4 *>(reinterpret_cast<const bf16 *>(d.in) + (size_t)c * 8), v);
      if (d.pro == PP_RMSNORM) {
        float ss = 0.f;
        if (mine) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
        }
        ss = block_sum2(ss, sc, nw, lane, w);
        const float inv = rsqrtf(ss * (1.0f / (float)d.K) + a.eps);
        if (mine) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = rbf(__fmul_rn(rbf(__fmul_rn(v[j], inv)), g[j]));
        }
      }
      if (mine) store_chunk_xs(xs, c, v);
    }
    __syncthreads();
    if (192 + ph < 512) tl_stamp(tp, 1);

    // ---- (B) this warp's units ------------------------------------------------------------------------------------------
    int pen_id = -1;
    if (d.epi == PE_FASTLOGITS && use_pen && lane < DA_WIN) pen_id = st->win[(d.pos + 1) * DA_WIN + lane];
    auto epilogue = [&](int j, float s0, float s1) {       // j: pair slot of this CTA; s0, s1: lane partial sums
      const int p = j * grid + bid, r0 = 2 * p, r1 = r0 + 1;
      const bool has1 = r1 < d.rows;
      // residual values first: their L2 round trip overlaps the butterflies
      float q0 = 0.f, q1 = 0.f;
      if (d.epi == PE_RESIDUAL && lane == 0) {
        if (d.res_units) { const uint32_t *ru = reinterpret_cast<const uint32_t *>(d.res); q0 = unit_val(ld_poll1(ru + r0)); if (has1) q1 = unit_val(ld_poll1(ru + r1)); }
        else { const bf16 *rp = reinterpret_cast<const bf16 *>(d.res); q0 = bf2f(rp[r0]); if (has1) q1 = bf2f(rp[r1]); }
      }
      float d0 = warp_sum(s0), d1 = warp_sum(s1);
      if (d.bias) { d0 += bf2f(d.bias[r0]); if (has1) d1 += bf2f(d.bias[r1]); }
      if (d.epi == PE_STORE) {
        if (lane == 0) { st_unit(d.out + r0, make_unit(d0, tag)); if (has1) st_unit(d.out + r1, make_unit(d1, tag)); }
      } else if (d.epi == PE_RESIDUAL) {
        if (lane == 0) { st_unit(d.out + r0, make_unit(q0 + rbf(d0), tag)); if (has1) st_unit(d.out + r1, make_unit(q1 + rbf(d1), tag)); }
      } else if (d.epi == PE_SWIGLU) {
        if (lane == 0) {
          const float gg = rbf(d0), up = rbf(d1);
          const float sg = rbf(gg / (1.0f + expf(-gg)));
          st_unit(d.out + p, make_unit(__fmul_rn(sg, up), tag));
        }
      } else {
        float z0 = rbf(d0), z1 = rbf(d1);
        const size_t lo = (size_t)(d.pos - 1) * a.fv;
        if (lane == 0) { a.flogits_raw[lo + r0] = f2bf(z0); if (has1) a.flogits_raw[lo + r1] = f2bf(z1); }
        const unsigned hit0 = __ballot_sync(0xffffffffu, pen_id == r0), hit1 = __ballot_sync(0xffffffffu, pen_id == r1);
        if (hit0) z0 = penalise(z0, rp_eff);
        if (hit1) z1 = penalise(z1, rp_eff);
        if (lane == 0) {
          a.flogits[lo + r0] = f2bf(z0); st_unit(d.out + r0, make_unit(z0, tag));
          if (has1) { a.flogits[lo + r1] = f2bf(z1); st_unit(d.out + r1, make_unit(z1, tag)); }
        }
      }
    };
    while (c_valid && c_ph == ph) {
      const int u = c_u, j = u / nb, b = u - j * nb;
      float p0 = 0.f, p1 = 0.f;
      fma_batch(ra, xs, b, nseg, lane, p0, p1);
      // the registers are free again: put this warp's NEXT unit in flight (often a later phase)
      c_valid = seek(c_ph, c_u + nw);
      if (c_valid) issue(ra, c_ph, c_u);
      if (!split) epilogue(j, p0, p1);
      else if (u < DA_PART_UNITS) { part[u * 64 + lane] = p0; part[u * 64 + 32 + lane] = p1; }
    }
    if (split) {
      __syncthreads();
      for (int j = w; j < npc; j += nw) {
        float s0 = 0.f, s1 = 0.f;
        for (int b = 0; b < nb; ++b) { s0 += part[(j * nb + b) * 64 + lane]; s1 += part[(j * nb + b) * 64 + 32 + lane]; }
        epilogue(j, s0, s1);
      }
    }
    if (192 + ph < 512) tl_stamp(tp, 2);

    // ---- (C) fast head: warps 0-3 of CTA 0 draw the code and publish its embedding as the next pass's input ---------------------
    if (d.epi == PE_FASTLOGITS && bid == 0) {
      __shared__ uint32_t s_tok;
      if (w < 4) {
        const int V = a.fv;
        if (threadIdx.x == 0) ok = wait_sentinel(d.out + V - 1, tag) && ok;
        named_bar(1, 128);
        uint32_t key[DA_G_IPT], idx[DA_G_IPT], valid = 0;
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < DA_G_IPT; ++i) {
          const int e = i * 128 + threadIdx.x;              // coalesced
          idx[i] = (uint32_t)e; key[i] = 0;
          if (e < V) {
            uint32_t uu; int it = 0;
            do { uu = ld_poll1(d.out + e); if ((uu & 0xFFFFu) == tag) break; __nanosleep(100); } while (++it < DA_SPIN_LIMIT);
            ok = ok && it < DA_SPIN_LIMIT;
            key[i] = bf16_key((uint16_t)(uu >> 16)); valid |= 1u << i; mx = fmaxf(mx, unit_val(uu));
          }
        }
        SampleParams sp;
        float *gs = reinterpret_cast<float *>(scr + 28);
        mx = warp_max(mx);
        if (lane == 0) gs[w] = mx;
        named_bar(1, 128);
        sp.m = fmaxf(fmaxf(gs[0], gs[1]), fmaxf(gs[2], gs[3]));
        {   // sum of exp terms as 2^-40 fixed point: order-free, identical to the per-phase path
          Red4 es = {0ull, 0, -1}; int par = 0;
#pragma unroll
          for (int i = 0; i < DA_G_IPT; ++i) if ((valid >> i) & 1u) es.s += (unsigned long long)(expf(bits2f(key_bf16(key[i])) - sp.m) * DA_FIX2_SCALE);
          named_bar(1, 128);
          sp.S = __ull2float_rn(group_reduce(es, scr, par, lane, w).s) * (1.0f / DA_FIX2_SCALE);
          named_bar(1, 128);
        }
        sp.T_bf = eff_temperature(st);
        sp.c_max = cmax_from_top_p(st->top_p);
        uint32_t tok = sample_group(key, idx, valid, (uint32_t)V, sp, st, (uint32_t)d.pos, a.noise_off0 + (long long)(d.pos - 1) * a.fv,
                                    &st->nucleus[d.pos], scr, lane, w);
        if (tok >= (uint32_t)a.codebook_size) { tok = a.codebook_size - 1; if (threadIdx.x == 0) st->err = 3; }
        if (threadIdx.x == 0) { s_tok = tok; st->tok_out[d.pos + 1] = (int)tok; }
      }
      __syncthreads();
      if (d.pos < a.ncb - 1) {
        const uint32_t tok = s_tok;
        for (int dd = threadIdx.x; dd < a.dim; dd += DA_P_THREADS) st_unit(a.u_fin + dd, make_unit(bf2f(a.fast_emb[(size_t)tok * a.dim + dd]), tag));
      }
    }
  }

  tl_stamp(a.tl, 3);
  if (!ok && threadIdx.x == 0) st->err = 4;
  if (bid == 0) {
    __syncthreads();
    if (threadIdx.x == 0) {
      st->phase_ctr = tag_base + (unsigned)nph;
      GemvArgs g; g.st = st; g.seq = a.seq; g.seq_stride = a.seq_stride; g.im_end_id = a.im_end_id; g.n_rows_tok = a.n_rows_tok;
      finish_step(g);
    }
  }
}

}  // namespace da
