// sampler.cuh -- repetition penalty / top-p / temperature / Exp(1)-race sampling, sort-free for the
// 155,776-way slow head.  Restates fish_tts/models/inference.py:24-80 with the reference's own
// rounding points (the whole sampler is bf16 there):
//   score<0 ? score*rp : score/rp      (rp rounded to bf16, result rounded to bf16)      :38-45
//   p_i   = bf16(exp(z_i - max) / sum)                     softmax over the sorted logits :48-51
//   cum_j = running sum of p in sorted order; remove_j = bf16(cum_j) > bf16(top_p); keep j=0 :49-53
//   z'_i  = bf16(z_i / bf16(clip(T,1e-5))) for kept, -inf otherwise                        :57-58
//   p'_i  = bf16(exp(z'_i - max') / sum');  r_i = bf16(p'_i / q_i);  argmax (first index on ties) :60, 26-27
//
// Where the reference is implementation-defined we fix a definition (DESIGN.md "sampler"):
//   * equal logits are ordered by ascending index (torch.sort leaves it unspecified);
//   * the running sum is exact: probabilities are bf16 values, summed as 2^-44 fixed point in
//     u64, so the nucleus does not depend on summation order (the reference's CPU path runs a
//     sequential fp32 sum, its CUDA path a bf16 tree scan -- they already disagree with each other).
#pragma once
#include <limits.h>

#include "common.cuh"

namespace da {

#define DA_FIX_SCALE 17592186044416.0f   // 2^44

struct SampleParams {
  float m, S;        // max and sum(exp(z-m)) over the (penalised) logits
  float T_bf;        // bf16(clip(temperature, 1e-5)) as float
  unsigned long long c_max;  // largest fixed-point cumulative sum that still rounds to <= bf16(top_p)
};

__device__ __forceinline__ unsigned long long pweight(float z, float m, float S) {
  float p = rbf(expf(z - m) / S);
  return (unsigned long long)(p * DA_FIX_SCALE);
}

__device__ __forceinline__ unsigned long long cmax_from_top_p(float top_p) {
  // bf16(cum) <= t  <=>  cum < mid, or cum == mid when t's mantissa is even (round-half-even)
  uint16_t tb = f2bits(top_p);
  float t = bits2f(tb), nxt = bits2f((uint16_t)(tb + 1));
  float mid = 0.5f * t + 0.5f * nxt;             // exact: 9 significant bits
  unsigned long long M = (unsigned long long)(mid * DA_FIX_SCALE);
  return (tb & 1u) ? (M ? M - 1 : 0) : M;
}

__device__ __forceinline__ float penalise(float z, float rp_bf) {
  return z < 0.f ? rbf(z * rp_bf) : rbf(z / rp_bf);
}
// torch keeps a 0-dim fp32 operand of a bf16 op in fp32 on the CPU but casts it to bf16 inside CUDA kernels
// (TensorIterator dynamic casting): `logits / temperature`, `score * repetition_penalty` (inference.py:42-44, 58)
__device__ __forceinline__ float eff_rep_penalty(const DAState *st) { return st->cpu_sem ? st->rep_penalty : rbf(st->rep_penalty); }
__device__ __forceinline__ float eff_temperature(const DAState *st) {
  float t = fmaxf(st->temperature, 1e-5f);
  return st->cpu_sem ? t : rbf(t);
}

// where the Exp(1) draws come from: an explicit bf16 block (tests, "shared seeded Philox RNG") or Philox keyed by
// (seed, step, head, element).  Loaded from the request state ONCE: the state lives in global memory and every access
// is an L2 round trip on the decode critical path.
struct NoiseSrc { const bf16 *noise; unsigned long long seed; unsigned int step; };
__device__ __forceinline__ NoiseSrc noise_src(const DAState *st) { NoiseSrc n = {st->noise, st->seed, st->step_ctr}; return n; }
__device__ __forceinline__ float noise_at(const NoiseSrc &ns, uint32_t head, long long head_off, uint32_t idx) {
  if (ns.noise) return bf2f(ns.noise[head_off + idx]);
  return exp1_noise(ns.seed, ns.step, head, idx);
}

__device__ __forceinline__ unsigned long long make_sortkey(uint16_t zbits, uint32_t idx) {
  return ((unsigned long long)(0xFFFFu - bf16_key(zbits)) << 32) | idx;   // ascending = logit desc, idx asc
}
__device__ __forceinline__ float sortkey_logit(unsigned long long k) {
  return bits2f(key_bf16(0xFFFFu - (uint32_t)(k >> 32)));
}

// block-wide exclusive scan of one u64 per thread (<= 1024 threads); scratch >= 33 u64
template <class B = BlockAll>
__device__ __forceinline__ unsigned long long block_excl_scan_u64(unsigned long long v, unsigned long long *scratch,
                                                                  unsigned long long *total) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = B::nthreads() >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  B::sync();
  if (lane == 31) scratch[w] = inc;
  B::sync();
  if (w == 0) {
    unsigned long long x = lane < nw ? scratch[lane] : 0ull, xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += t;
    }
    scratch[lane] = xi - x;            // exclusive warp offsets
    if (lane == 31) scratch[32] = xi;  // grand total
  }
  B::sync();
  if (total) *total = scratch[32];
  return scratch[w] + inc - v;
}

struct ArgBest { float r; uint32_t idx; };
__device__ __forceinline__ ArgBest better(ArgBest a, ArgBest b) {   // larger r, then smaller idx
  if (b.r > a.r || (b.r == a.r && b.idx < a.idx)) return b;
  return a;
}
template <class B = BlockAll>
__device__ __forceinline__ ArgBest block_argbest(ArgBest v, float *fs, uint32_t *is) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = B::nthreads() >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgBest t; t.r = __shfl_xor_sync(0xffffffffu, v.r, o); t.idx = __shfl_xor_sync(0xffffffffu, v.idx, o);
    v = better(v, t);
  }
  B::sync();
  if (lane == 0) { fs[w] = v.r; is[w] = v.idx; }
  B::sync();
  if (threadIdx.x == 0) {
    ArgBest b = {fs[0], is[0]};
    for (int i = 1; i < nw; ++i) { ArgBest t = {fs[i], is[i]}; b = better(b, t); }
    fs[32] = b.r; is[32] = b.idx;
  }
  B::sync();
  ArgBest out = {fs[32], is[32]};
  return out;
}

// ---- block reduction of (u64 sum, int sum, int max) with ONE barrier per call ------------------------
// scratch: 2 x 3 x 32 u64 (ping-pong by call parity, see DESIGN.md "sampler"); all threads get the result
struct Red { unsigned long long s; int c; int m; };
template <class B = BlockAll>
__device__ __forceinline__ Red block_reduce(Red v, unsigned long long *scr, int &parity) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = B::nthreads() >> 5;
  // warp stage with redux.sync: the u64 sum as three 22-bit limbs (32 lanes x 2^22 < 2^32), count and max directly --
  // five single-step reductions instead of five dependent shuffle rounds over four registers
  const unsigned l0 = __reduce_add_sync(0xffffffffu, (unsigned)v.s & 0x3FFFFFu);
  const unsigned l1 = __reduce_add_sync(0xffffffffu, (unsigned)(v.s >> 22) & 0x3FFFFFu);
  const unsigned l2 = __reduce_add_sync(0xffffffffu, (unsigned)(v.s >> 44));
  v.s = (unsigned long long)l0 + ((unsigned long long)l1 << 22) + ((unsigned long long)l2 << 44);
  v.c = __reduce_add_sync(0xffffffffu, v.c);
  v.m = __reduce_max_sync(0xffffffffu, v.m);
  unsigned long long *b = scr + parity * 96;      // [nw] sums, then [nw] (count << 32 | max)
  parity ^= 1;
  if (lane == 0) { b[w] = v.s; b[32 + w] = ((unsigned long long)(unsigned)v.c << 32) | (unsigned)v.m; }
  B::sync();
  Red r = {0ull, 0, INT_MIN};
  for (int i = 0; i < nw; ++i) { const unsigned long long y = b[32 + i]; r.s += b[i]; r.c += (int)(y >> 32); r.m = max(r.m, (int)(unsigned)y); }
  return r;
}

#define DA_FIX2_SCALE 1099511627776.0f   // 2^40: fixed point of the second softmax's exp terms

// ---- selection over register-resident items: no sort ----------------------------------------------------
// Each thread holds IPT items (16-bit monotone logit key, vocabulary index).  The nucleus {j : cum_j <= c_max} of the
// (logit desc, index asc) order is found by bisection on the key (all sums exact u64 => order-free), a partially
// kept tie group by a second bisection on the index.  Returns the sampled index, or 0xFFFFFFFF when the nucleus is
// not proven to lie inside the item set (caller falls back); `all_present`: the items are the whole vocabulary.
template <int IPT, class B = BlockAll>
__device__ uint32_t sample_items(const uint32_t (&key)[IPT], const uint32_t (&idx)[IPT], uint32_t valid_mask, uint32_t idx_limit,
                                 bool all_present, const SampleParams &sp, const NoiseSrc &st, uint32_t head, long long head_off,
                                 int *nucleus_out, unsigned long long *scr) {
  int parity = 0;
  unsigned long long w[IPT];
  Red t = {0ull, 0, -1};
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const bool v = (valid_mask >> i) & 1u;
    w[i] = v ? pweight(bits2f(key_bf16(key[i])), sp.m, sp.S) : 0ull;
    t.s += w[i]; t.c += v; if (v) t.m = max(t.m, (int)key[i]);
  }
  const Red tot = block_reduce<B>(t, scr, parity);      // total mass, item count, top key
  const int n = tot.c, top_key = tot.m;
  // (1) lowest key kappa with G(kappa) = sum_{key >= kappa} w <= c_max
  uint32_t lo = 0, hi = 65536;
  if (tot.s <= sp.c_max) hi = 0;
  else while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    Red g = {0ull, 0, -1};
#pragma unroll
    for (int i = 0; i < IPT; ++i) if (((valid_mask >> i) & 1u) && key[i] >= mid) g.s += w[i];
    if (block_reduce<B>(g, scr, parity).s <= sp.c_max) hi = mid; else lo = mid;
  }
  const uint32_t kappa = hi;
  // (2) mass and count of the fully kept groups, next lower present key
  Red g = {0ull, 0, -1};
#pragma unroll
  for (int i = 0; i < IPT; ++i) if ((valid_mask >> i) & 1u) {
    if (key[i] >= kappa) { g.s += w[i]; g.c += 1; } else g.m = max(g.m, (int)key[i]);
  }
  g = block_reduce<B>(g, scr, parity);
  const int n_full = g.c, tau = g.m;
  // (3) partially kept tie group
  int c_part = 0; long long i_cut = -1;
  if (tau >= 0) {
    Red q = {0ull, 0, -1};
#pragma unroll
    for (int i = 0; i < IPT; ++i) if (((valid_mask >> i) & 1u) && (int)key[i] == tau) q.c += 1;
    const int n_tau = block_reduce<B>(q, scr, parity).c;
    const unsigned long long wt = pweight(bits2f(key_bf16((uint32_t)tau)), sp.m, sp.S);
    const unsigned long long room = sp.c_max >= g.s ? sp.c_max - g.s : 0ull;
    const unsigned long long c = wt ? room / wt : (unsigned long long)n_tau;
    c_part = (int)(c < (unsigned long long)n_tau ? c : (unsigned long long)n_tau);
    if (n_full == 0 && c_part < 1) c_part = 1;        // always keep the top-1 (inference.py:53)
    if (c_part == n_tau) i_cut = (long long)idx_limit;
    else if (c_part > 0) {                             // index of the c_part-th member in index order
      long long l = -1, h = (long long)idx_limit;
      while (h - l > 1) {
        const long long mid = (l + h) >> 1;
        Red z = {0ull, 0, -1};
#pragma unroll
        for (int i = 0; i < IPT; ++i) if (((valid_mask >> i) & 1u) && (int)key[i] == tau && (long long)idx[i] <= mid) z.c += 1;
        if (block_reduce<B>(z, scr, parity).c >= c_part) h = mid; else l = mid;
      }
      i_cut = h;
    }
  }
  const int n_keep = n_full + c_part;
  if (n_keep == n && !all_present) return 0xFFFFFFFFu;
  if (threadIdx.x == 0 && nucleus_out) *nucleus_out = n_keep;
  // (4) second softmax over the kept set (exp terms summed as 2^-40 fixed point => order-free), then the Exp(1) race
  const float mz = rbf(bits2f(key_bf16((uint32_t)top_key)) / sp.T_bf);
  float e2[IPT]; uint32_t keep = 0;
  Red s2 = {0ull, 0, -1};
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const bool k = ((valid_mask >> i) & 1u) && (key[i] >= kappa || ((int)key[i] == tau && (long long)idx[i] <= i_cut));
    e2[i] = 0.f;
    if (k) { keep |= 1u << i; e2[i] = expf(rbf(bits2f(key_bf16(key[i])) / sp.T_bf) - mz); s2.s += (unsigned long long)(e2[i] * DA_FIX2_SCALE); }
  }
  const float S2 = __ull2float_rn(block_reduce<B>(s2, scr, parity).s) * (1.0f / DA_FIX2_SCALE);
  ArgBest best = {0.f, 0u};   // removed tokens have probability 0 -> r = 0; argmax ties go to index 0
#pragma unroll
  for (int i = 0; i < IPT; ++i) if ((keep >> i) & 1u) {
    const float p2 = rbf(e2[i] / S2);
    ArgBest cnd = {rbf(p2 / noise_at(st, head, head_off, idx[i])), idx[i]};
    best = better(best, cnd);
  }
  B::sync();
  float *fs = reinterpret_cast<float *>(scr);
  best = block_argbest<B>(best, fs, reinterpret_cast<uint32_t *>(fs + 40));
  return best.idx;
}


// ---- sort-based nucleus sampler (the persistent kernel's fast path) -------------------------------------------------
// The definition above orders the items by (logit desc, index asc), keeps the longest prefix whose exact cumulative
// weight is <= c_max (and always the first item), renormalises and runs the Exp(1) race.  sample_items finds that prefix
// without sorting by ~30 block-wide bisection rounds; with a handful of warps each round costs a barrier, which put the
// nine fast heads at ~25 us each on the decode critical path.  Here the NT*E items are SORTED by a bitonic network
// (register / shuffle / shared-memory stages), the prefix falls out of one scan, and the whole head costs ~2 us.  All sums
// are the same exact u64 fixed-point sums, so both routines return the same token for the same logits, bit for bit.
//
// Items: 32-bit composites ((0xFFFF - key16) << 16) | tie, ascending = (logit desc, tie asc); `tie` is the vocabulary
// index (fast heads) or a slot number that increases with the vocabulary index (slow-head candidate list; slot2idx maps
// back).  Padding items are 0xFFFFFFFF.  Thread t of the NT-thread group holds items t*E .. t*E+E-1.
// The k / j loops are RUNTIME loops on purpose: a fully unrolled network is thousands of straight-line instructions
// that execute once per head, i.e. at instruction-fetch speed (measured: the unrolled sampler took 17-30 us per head).
// The items stay in registers: only the in-thread strides need compile-time indices (three small blocks selected by j),
// and for strides >= E the min/max direction is the same for all items of a thread, so a stage is E shuffles + E min/max.
template <int E, int J>
__device__ __forceinline__ void bitonic_inthread(uint32_t (&a)[E], bool asc) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if ((e & J) == 0 && (e | J) < E) {
      const uint32_t lo = min(a[e], a[e | J]), hi = max(a[e], a[e | J]);
      a[e] = asc ? lo : hi; a[e | J] = asc ? hi : lo;
    }
  }
}
// in-thread tail of a merge step k >= E: strides E/2 .. 1, one direction for the whole thread
template <int E>
__device__ __forceinline__ void bitonic_merge_inthread(uint32_t (&a)[E], bool asc) {
  if (E > 8) bitonic_inthread<E, 8>(a, asc);
  if (E > 4) bitonic_inthread<E, 4>(a, asc);
  if (E > 2) bitonic_inthread<E, 2>(a, asc);
  if (E > 1) bitonic_inthread<E, 1>(a, asc);
}
template <int E, int NT, class B>
__device__ __forceinline__ void bitonic_sort_u32(uint32_t (&a)[E], uint32_t *sm) {
  static_assert(E == 2 || E == 4 || E == 8 || E == 16, "items per thread");
  const int t = threadIdx.x;
  // (1) levels k < E: sort inside the thread (static network, directions alternate with the element index)
#pragma unroll
  for (int k = 2; k < E; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if ((e & j) == 0) {
          const bool asc = (e & k) == 0;
          const uint32_t lo = min(a[e], a[e | j]), hi = max(a[e], a[e | j]);
          a[e] = asc ? lo : hi; a[e | j] = asc ? hi : lo;
        }
      }
    }
  }
  // (2) levels k >= E: the direction is a property of the thread; strides >= E cross lanes / warps, then one in-thread tail
#pragma unroll 1
  for (int k = E; k <= NT * E; k <<= 1) {
    const bool asc = ((t * E) & k) == 0;
#pragma unroll 1
    for (int j = k >> 1; j >= E; j >>= 1) {
      const bool takemin = (((t * E) & j) == 0) == asc;
      if (j >= 32 * E) {            // partner lives in another warp: through shared memory
#pragma unroll
        for (int e = 0; e < E; ++e) sm[t * E + e] = a[e];
        B::sync();
        const uint32_t *pp = sm + ((t * E) ^ j);
#pragma unroll
        for (int e = 0; e < E; ++e) { const uint32_t p = pp[e]; a[e] = takemin ? min(a[e], p) : max(a[e], p); }
        B::sync();
      } else {                      // partner is another lane of this warp
        const int lx = j / E;
#pragma unroll
        for (int e = 0; e < E; ++e) { const uint32_t p = __shfl_xor_sync(0xffffffffu, a[e], lx); a[e] = takemin ? min(a[e], p) : max(a[e], p); }
      }
    }
    bitonic_merge_inthread<E>(a, asc);
  }
}

// scr: >= 2 * 3 * 32 u64 (block_reduce layout) ; sm_sort: NT*E u32.  n: number of real items.  all_present: the items are
// the whole vocabulary.  Returns the sampled vocabulary index, or 0xFFFFFFFF when the nucleus is not proven to lie inside
// the item set (caller falls back).
#ifdef DA_SAMPLER_TIMING
__device__ long long g_sampler_t[16];
#define DA_ST(i) do { if (threadIdx.x == 0) g_sampler_t[i] = clock64(); } while (0)
#else
#define DA_ST(i) do { } while (0)
#endif
template <int E, int NT, class B>
__device__ __noinline__ uint32_t sample_sorted(uint32_t (&a_in)[E], uint32_t n, bool all_present, const unsigned long long *slot2idx, const SampleParams &sp,
                                  const NoiseSrc &st, uint32_t head, long long head_off, int *nucleus_out, uint32_t *sm_sort, unsigned long long *scr) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  constexpr int NW = NT / 32;
  uint32_t a[E];                                   // registers (the argument lives in the caller's frame)
#pragma unroll
  for (int e = 0; e < E; ++e) a[e] = a_in[e];
  DA_ST(0);
  bitonic_sort_u32<E, NT, B>(a, sm_sort);
  DA_ST(1);
  // exact inclusive cumulative weight in sorted order (exp and division of the E items are independent: issue them together)
  float ez[E];
#pragma unroll
  for (int e = 0; e < E; ++e) ez[e] = expf(bits2f(key_bf16(0xFFFFu - (a[e] >> 16))) - sp.m);
  unsigned long long cum[E];
  unsigned long long run = 0ull;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool v = a[e] != 0xFFFFFFFFu;
    run += v ? (unsigned long long)(rbf(ez[e] / sp.S) * DA_FIX_SCALE) : 0ull;      // = pweight()
    cum[e] = run;
  }
  DA_ST(2);
  unsigned long long inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += x;
  }
  unsigned long long *wsum = scr;                 // [NW] warp totals
  int *wcnt = reinterpret_cast<int *>(scr + 40);  // [NW] kept items per warp (second round)
  if (lane == 31) wsum[w] = inc;
  // the sorted items go to shared memory: the kept prefix is redistributed over all threads below (after the sort it sits in
  // the first few threads, which would otherwise run the Philox draws of the race one after the other)
#pragma unroll
  for (int e = 0; e < E; ++e) sm_sort[t * E + e] = a[e];
  B::sync();
  unsigned long long base = inc - run;
#pragma unroll
  for (int i = 0; i < NW; ++i) if (i < w) base += wsum[i];
  int kept = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) kept += (a[e] != 0xFFFFFFFFu) && (base + cum[e] <= sp.c_max || (t == 0 && e == 0));
  kept = (int)__reduce_add_sync(0xffffffffu, (unsigned)kept);
  if (lane == 0) wcnt[w] = kept;
  B::sync();
  int n_keep = 0;
#pragma unroll
  for (int i = 0; i < NW; ++i) n_keep += wcnt[i];       // the kept set is a prefix of the sorted order
  DA_ST(3);
  if (n_keep == (int)n && !all_present) return 0xFFFFFFFFu;
  if (t == 0 && nucleus_out) *nucleus_out = n_keep;
  // second softmax over the kept prefix (exp terms summed as 2^-40 fixed point => order-free), items dealt round-robin
  const uint32_t top = sm_sort[0];
  const float mz = rbf(bits2f(key_bf16(0xFFFFu - (top >> 16))) / sp.T_bf);
  float e2[E]; uint32_t it[E];
  Red s2 = {0ull, 0, -1};
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = t + e * NT;
    e2[e] = 0.f; it[e] = 0xFFFFFFFFu;
    if (i < n_keep) {
      it[e] = sm_sort[i];
      e2[e] = expf(rbf(bits2f(key_bf16(0xFFFFu - (it[e] >> 16))) / sp.T_bf) - mz);
      s2.s += (unsigned long long)(e2[e] * DA_FIX2_SCALE);
    }
  }
  DA_ST(4);
  int parity = 0;
  B::sync();                                       // wsum / wcnt are dead: scr becomes block_reduce scratch
  s2 = block_reduce<B>(s2, scr, parity);
  const float S2 = __ull2float_rn(s2.s) * (1.0f / DA_FIX2_SCALE);
  DA_ST(5);
  ArgBest best = {0.f, 0u};   // removed tokens have probability 0 -> r = 0; argmax ties go to index 0
#pragma unroll
  for (int e = 0; e < E; ++e) if (it[e] != 0xFFFFFFFFu) {
    const uint32_t tie = it[e] & 0xFFFFu;
    const uint32_t idx = slot2idx ? (uint32_t)((slot2idx[tie] >> 30) & 0x3FFFFu) : tie;      // candidate entry: key (16) | index (18) | tag (30)
    const float p2 = rbf(e2[e] / S2);
    ArgBest cnd = {rbf(p2 / noise_at(st, head, head_off, idx)), idx};
    best = better(best, cnd);
  }
  DA_ST(6);
  B::sync();
  float *fs = reinterpret_cast<float *>(scr);
  best = block_argbest<B>(best, fs, reinterpret_cast<uint32_t *>(fs + 40));
  DA_ST(7);
  return best.idx;
}

// ---- binned nucleus sampler -----------------------------------------------------------------------
// Same result as sample_sorted(), without sorting.  The nucleus is a prefix of the (logit desc, index asc) order, so only
// the position of the cut has to be exact.  Items are binned by their distance below the maximum (NB/8 bins per unit;
// everything further than 8 below shares the last bin) -- monotone in the logit, so bins respect the order.  Per bin the
// kernel accumulates the exact fixed-point weight (three 15-bit limbs, 32-bit shared atomics: integer, order-free) and
// the item count; one block scan finds the bin whose cumulative weight first exceeds c_max.  Bins ahead of it are kept
// whole, bins after it dropped whole, and only the items of the cut bin (a few dozen) are ranked exactly against each
// other.  The kept items stay with the threads that hold them, so the Philox draws of the race run in parallel.
// sm: 4*NB + 2*NT + 8 words + NT u64 (and >= NT*E words for the sort fallback, taken when the cut bin has > NT items).
template <int E, int NT, class B, int NB = 512>
__device__ __noinline__ uint32_t sample_binned(uint32_t (&a_in)[E], uint32_t n, bool all_present, const unsigned long long *slot2idx, const SampleParams &sp,
                                  const NoiseSrc &st, uint32_t head, long long head_off, int *nucleus_out, uint32_t *sm, unsigned long long *scr) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  constexpr int NW = NT / 32, BPT = (NB + NT - 1) / NT;
  uint32_t *h0 = sm, *h1 = sm + NB, *h2 = sm + 2 * NB, *hc = sm + 3 * NB;
  uint32_t *list = sm + 4 * NB, *flag = list + NT, *misc = flag + NT;
  unsigned long long *wl = reinterpret_cast<unsigned long long *>(misc + 8);
  for (int i = t; i < 4 * NB; i += NT) sm[i] = 0u;
  if (t == 0) { misc[0] = 0u; misc[1] = (uint32_t)NB; misc[2] = 0u; misc[3] = 0u; misc[4] = 0u; }
  B::sync();
  DA_ST(0);
  uint32_t a[E]; int bin[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    a[e] = a_in[e];
    bin[e] = NB;
    if (a[e] != 0xFFFFFFFFu) {
      const float z = bits2f(key_bf16(0xFFFFu - (a[e] >> 16)));
      const unsigned long long wv = pweight(z, sp.m, sp.S);
      const int b = (int)fminf(fmaxf((sp.m - z) * (float)(NB / 8), 0.f), (float)(NB - 1));      // clamps keep the map monotone
      bin[e] = b;
      const uint32_t l0 = (uint32_t)wv & 0x7FFFu, l1 = (uint32_t)(wv >> 15) & 0x7FFFu, l2 = (uint32_t)(wv >> 30);
      if (l0) atomicAdd(h0 + b, l0);
      if (l1) atomicAdd(h1 + b, l1);
      if (l2) atomicAdd(h2 + b, l2);
      atomicAdd(hc + b, 1u);
    }
  }
  B::sync();
  DA_ST(1);
  // block scan over the bins (thread t owns bins t*BPT ..): inclusive weight and count
  unsigned long long hs[BPT], bw[BPT]; int cs[BPT];
  unsigned long long run = 0ull; int crun = 0;
#pragma unroll
  for (int i = 0; i < BPT; ++i) {
    const int b = t * BPT + i;
    bw[i] = 0ull;
    if (b < NB) { bw[i] = (unsigned long long)h0[b] + ((unsigned long long)h1[b] << 15) + ((unsigned long long)h2[b] << 30); crun += (int)hc[b]; }
    run += bw[i]; hs[i] = run; cs[i] = crun;
  }
  unsigned long long inc = run; int cinc = crun;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long x = __shfl_up_sync(0xffffffffu, inc, o);
    const int y = __shfl_up_sync(0xffffffffu, cinc, o);
    if (lane >= o) { inc += x; cinc += y; }
  }
  unsigned long long *wsum = scr + 192;               // [NW] warp totals; clear of both block_reduce buffers (scr[0..48), scr[96..144))
  int *wcnt = reinterpret_cast<int *>(scr + 208);     // [NW]
  if (lane == 31) { wsum[w] = inc; wcnt[w] = cinc; }
  B::sync();
  unsigned long long base = inc - run; int cbase = cinc - crun;
#pragma unroll
  for (int i = 0; i < NW; ++i) if (i < w) { base += wsum[i]; cbase += wcnt[i]; }
#pragma unroll
  for (int i = 0; i < BPT; ++i) {
    const unsigned long long incl = base + hs[i], excl = incl - bw[i];
    if (excl <= sp.c_max && incl > sp.c_max) {      // the one bin where the cumulative weight crosses c_max
      const int b = t * BPT + i;
      misc[1] = (uint32_t)b; misc[2] = (uint32_t)excl; misc[3] = (uint32_t)(excl >> 32); misc[4] = (uint32_t)(cbase + cs[i] - (int)hc[b]);
    }
  }
  B::sync();
  const int cut = (int)misc[1];
  const unsigned long long before = (unsigned long long)misc[2] | ((unsigned long long)misc[3] << 32);
  const int cnt_before = (int)misc[4];
  DA_ST(2);
  int mypos[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    mypos[e] = -1;
    if (a[e] != 0xFFFFFFFFu && bin[e] == cut) {
      const int p = (int)atomicAdd(misc, 1u);
      mypos[e] = p;
      if (p < NT) { list[p] = a[e]; wl[p] = pweight(bits2f(key_bf16(0xFFFFu - (a[e] >> 16))), sp.m, sp.S); }
    }
  }
  B::sync();
  const int n2 = (int)misc[0];
  DA_ST(3);
  if (n2 > NT) {                                    // degenerate (a flat tail in one bin): the sorting sampler handles it
    B::sync();
    return sample_sorted<E, NT, B>(a_in, n, all_present, slot2idx, sp, st, head, head_off, nucleus_out, sm, scr);
  }
  if (t < n2) {
    const uint32_t ci = list[t];
    unsigned long long G = before; bool pre = false;
    for (int j = 0; j < n2; ++j) if (list[j] < ci) { G += wl[j]; pre = true; }
    flag[t] = ((G + wl[t] <= sp.c_max) || (!pre && cnt_before == 0)) ? 1u : 0u;      // the first item of the order is always kept
  }
  B::sync();
  DA_ST(4);
  // second softmax over the kept items (exp terms summed as 2^-40 fixed point => order-free) and the race
  const float mz = rbf(sp.m / sp.T_bf);             // the top item is the maximum
  float e2[E];
  uint32_t keep = 0u;
  Red s2 = {0ull, 0, -1};
#pragma unroll
  for (int e = 0; e < E; ++e) {
    e2[e] = 0.f;
    const bool k = a[e] != 0xFFFFFFFFu && (bin[e] < cut || (bin[e] == cut && flag[mypos[e]] != 0u));
    if (k) {
      keep |= 1u << e; ++s2.c;
      e2[e] = expf(rbf(bits2f(key_bf16(0xFFFFu - (a[e] >> 16))) / sp.T_bf) - mz);
      s2.s += (unsigned long long)(e2[e] * DA_FIX2_SCALE);
    }
  }
  int parity = 0;
  s2 = block_reduce<B>(s2, scr, parity);
  DA_ST(5);
  if (s2.c == (int)n && !all_present) return 0xFFFFFFFFu;
  if (t == 0 && nucleus_out) *nucleus_out = s2.c;
  const float S2 = __ull2float_rn(s2.s) * (1.0f / DA_FIX2_SCALE);
  ArgBest best = {0.f, 0u};   // removed tokens have probability 0 -> r = 0; argmax ties go to index 0
#pragma unroll
  for (int e = 0; e < E; ++e) if (keep & (1u << e)) {
    const uint32_t tie = a[e] & 0xFFFFu;
    const uint32_t idx = slot2idx ? (uint32_t)((slot2idx[tie] >> 30) & 0x3FFFFu) : tie;      // candidate entry: key (16) | index (18) | tag (30)
    const float p2 = rbf(e2[e] / S2);
    ArgBest cnd = {rbf(p2 / noise_at(st, head, head_off, idx)), idx};
    best = better(best, cnd);
  }
  DA_ST(6);
  // winner: larger r, then smaller index.  r >= +0, so its bits order like the value: one u64 key per warp from two
  // redux steps, one barrier, every thread scans the NW keys (scr[96..): the reduction above used the other buffer)
  {
    const unsigned rb = __float_as_uint(best.r);
    const unsigned mx = __reduce_max_sync(0xffffffffu, rb);
    const unsigned mi = __reduce_min_sync(0xffffffffu, rb == mx ? best.idx : 0xFFFFFFFFu);
    unsigned long long *wk = scr + 96;
    if (lane == 0) wk[w] = ((unsigned long long)mx << 32) | (0xFFFFFFFFu - mi);
    B::sync();
    unsigned long long k = 0ull;
#pragma unroll
    for (int i = 0; i < NW; ++i) k = max(k, wk[i]);
    best.idx = 0xFFFFFFFFu - (uint32_t)k;
  }
  DA_ST(7);
  return best.idx;
}

// ---- fallback: nucleus wider than the candidate list (flat distributions) ------------------------
// One CTA walks the whole logits vector from global memory; all sums are u64, so order-free.
template <class B = BlockAll>
__device__ __noinline__ uint32_t sample_fallback(const bf16 *logits, int V, const SampleParams &sp, const NoiseSrc &st,
                                    uint32_t head, long long head_off, int *nucleus_out,
                                    unsigned long long *scr64, float *scrf) {
  const uint16_t *lb = reinterpret_cast<const uint16_t *>(logits);
  __shared__ int sh_i[4];
  // (1) lowest key kappa such that G(kappa) = sum_{key >= kappa} w <= c_max
  auto G_of = [&](uint32_t key) {
    unsigned long long g = 0;
    for (int i = threadIdx.x; i < V; i += B::nthreads()) {
      uint16_t b = lb[i];
      if (bf16_key(b) >= key) g += pweight(bits2f(b), sp.m, sp.S);
    }
    unsigned long long tot;
    block_excl_scan_u64<B>(g, scr64, &tot);
    return tot;
  };
  uint32_t lo = 0, hi = 65536;   // invariant: G(hi) <= c_max (G(65536) = 0), G(lo) > c_max
  if (G_of(0) <= sp.c_max) hi = 0;
  else while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (G_of(mid) <= sp.c_max) hi = mid; else lo = mid;
  }
  uint32_t kappa = hi;
  // (2) G(kappa), count of fully kept, and the next lower present key
  unsigned long long g = 0; int cnt = 0; uint32_t below = 0; bool has_below = false;
  for (int i = threadIdx.x; i < V; i += B::nthreads()) {
    uint16_t b = lb[i]; uint32_t k = bf16_key(b);
    if (k >= kappa) { g += pweight(bits2f(b), sp.m, sp.S); ++cnt; }
    else if (!has_below || k > below) { below = k; has_below = true; }
  }
  unsigned long long G;
  block_excl_scan_u64<B>(g, scr64, &G);
  int n_full = (int)(block_sum<B>((float)cnt, scrf) + 0.5f);
  float bmax = block_max<B>(has_below ? (float)below : -1.f, scrf);   // keys < 65536: exact in fp32
  int tau = (int)bmax;   // -1: nothing below
  // (3) partial group: how many members of key tau are kept, and up to which index
  int c_part = 0, i_cut = -1;
  if (tau >= 0) {
    int n_tau = 0;
    for (int i = threadIdx.x; i < V; i += B::nthreads()) n_tau += (bf16_key(lb[i]) == (uint32_t)tau);
    n_tau = (int)(block_sum<B>((float)n_tau, scrf) + 0.5f);
    unsigned long long w = pweight(bits2f(key_bf16((uint32_t)tau)), sp.m, sp.S);
    unsigned long long room = sp.c_max >= G ? sp.c_max - G : 0ull;
    unsigned long long c = w ? room / w : (unsigned long long)n_tau;
    c_part = (int)(c < (unsigned long long)n_tau ? c : (unsigned long long)n_tau);
    if (n_full == 0 && c_part < 1) c_part = 1;          // always keep the top-1 (inference.py:53)
    if (c_part > 0) {
      // index of the c_part-th member in index order: ordered block scan, chunk by chunk
      if (threadIdx.x == 0) { sh_i[0] = 0; sh_i[1] = -1; }
      B::sync();
      for (int base = 0; base < V; base += B::nthreads()) {
        int i = base + threadIdx.x;
        unsigned long long f = (i < V && bf16_key(lb[i]) == (uint32_t)tau) ? 1ull : 0ull;
        unsigned long long tot;
        unsigned long long ex = block_excl_scan_u64<B>(f, scr64, &tot);
        int before = sh_i[0];
        if (f && before + (int)ex + 1 == c_part) sh_i[1] = i;
        B::sync();
        if (threadIdx.x == 0) sh_i[0] = before + (int)tot;
        B::sync();
        if (sh_i[1] >= 0) break;
      }
      i_cut = sh_i[1];
    }
  }
  if (threadIdx.x == 0 && nucleus_out) *nucleus_out = n_full + c_part;
  // (4) second softmax + race over the kept set
  float mz = rbf(sp.m / sp.T_bf);   // the top logit is always kept
  unsigned long long es = 0;
  for (int i = threadIdx.x; i < V; i += B::nthreads()) {
    uint16_t b = lb[i]; uint32_t k = bf16_key(b);
    bool keep = k >= kappa || ((int)k == tau && i <= i_cut);
    if (keep) es += (unsigned long long)(expf(rbf(bits2f(b) / sp.T_bf) - mz) * DA_FIX2_SCALE);
  }
  unsigned long long es_tot;
  block_excl_scan_u64<B>(es, scr64, &es_tot);
  float S2 = __ull2float_rn(es_tot) * (1.0f / DA_FIX2_SCALE);
  ArgBest best = {0.f, 0u};
  for (int i = threadIdx.x; i < V; i += B::nthreads()) {
    uint16_t b = lb[i]; uint32_t k = bf16_key(b);
    bool keep = k >= kappa || ((int)k == tau && i <= i_cut);
    if (keep) {
      float p2 = rbf(expf(rbf(bits2f(b) / sp.T_bf) - mz) / S2);
      ArgBest c = {rbf(p2 / noise_at(st, head, head_off, (uint32_t)i)), (uint32_t)i};
      best = better(best, c);
    }
  }
  best = block_argbest<B>(best, scrf, (uint32_t *)(scrf + 40));
  return best.idx;
}

}  // namespace da
