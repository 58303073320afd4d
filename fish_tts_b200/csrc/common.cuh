// common.cuh -- shared device helpers for the dual-AR decode kernels (sm_100a).
//
// Numerics contract (SURVEY.md section 8a "Numerics summary"): every value the reference's eager
// bf16 path materialises as a tensor is rounded to bf16 here at the same point; everything in
// between is fp32.  The only freedom taken is the ORDER of fp32 sums, which the reference
// leaves to cuBLAS / ATen reductions.  No --use_fast_math: expf, division, rsqrtf are the same
// device functions torch's kernels call.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

#define DA_MAX_ROWS 16   // num_codebooks + 1 <= 16
#define DA_WIN 16        // repetition window width (inference.py:187)
#define DA_MAX_KV_HEADS 32
#define DA_CAND_CAP 8192 // candidate list capacity of the slow-head sampler

namespace da {

__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bits2f(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
__device__ __forceinline__ uint16_t f2bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }

// 8 packed bf16 -> 8 floats (exact)
__device__ __forceinline__ void unpack8(const uint4 &u, float *f) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

// ---- L2 cache policies (guide: "per-load cache hints are the more common tool") ---------------
// slow weights / LM head / KV stream through once per token -> evict_first;
// the fast stack is re-read num_codebooks times per token -> evict_last keeps it L2-resident.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
// 128-bit streaming weight load: read-only path, no L1 allocation, explicit L2 policy
__device__ __forceinline__ uint4 ldg_w(const void *p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}

// ---- warp / block reductions (fixed order => deterministic) -----------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// ---- block scope of a collective: the whole CTA (barrier 0) or a named barrier over the first N threads ----------
// The persistent decode kernel keeps a producer warp outside its 512 compute threads, so its block-wide
// reductions run on a named barrier; the per-phase kernels use the whole CTA.
struct BlockAll {
  static __device__ __forceinline__ void sync() { __syncthreads(); }
  static __device__ __forceinline__ int nthreads() { return blockDim.x; }
};
template <int ID, int N> struct BlockNamed {
  static __device__ __forceinline__ void sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
  static __device__ __forceinline__ int nthreads() { return N; }
};

// all threads get the result; `scratch` holds >= 33 floats; thread count multiple of 32
template <class B = BlockAll>
__device__ __forceinline__ float block_sum(float v, float *scratch) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = B::nthreads() >> 5;
  v = warp_sum(v);
  B::sync();
  if (lane == 0) scratch[w] = v;
  B::sync();
  if (w == 0) {
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += scratch[i];   // sequential, fixed order
    if (lane == 0) scratch[32] = t;
  }
  B::sync();
  return scratch[32];
}
template <class B = BlockAll>
__device__ __forceinline__ float block_max(float v, float *scratch) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = B::nthreads() >> 5;
  v = warp_max(v);
  B::sync();
  if (lane == 0) scratch[w] = v;
  B::sync();
  if (w == 0) {
    float t = -INFINITY;
    for (int i = 0; i < nw; ++i) t = fmaxf(t, scratch[i]);
    if (lane == 0) scratch[32] = t;
  }
  B::sync();
  return scratch[32];
}

// ---- activation vector staged in shared memory -------------------------------------------------
// The GEMV gives lane l of a warp elements [s*256 + l*8, +8) of segment s.  Stored as fp32 in
// two 16-byte halves so that a warp's LDS.128 is conflict-free:
//   element e -> ((s*2 + half)*32 + lane)*4 + j   with s=e/256, lane=(e%256)/8, half=(e%8)/4, j=e%4
__device__ __forceinline__ int xs_index(int e) {
  int s = e >> 8, r = e & 255, lane = r >> 3, h = (r >> 2) & 1, j = r & 3;
  return (((s << 1) + h) * 32 + lane) * 4 + j;
}

// ---- monotone 16-bit key of a bf16 value: larger value <=> larger key --------------------------
__device__ __forceinline__ uint32_t bf16_key(uint16_t b) {
  return (b & 0x8000u) ? (uint32_t)(0xFFFFu & ~b) : (uint32_t)(b | 0x8000u);
}
__device__ __forceinline__ uint16_t key_bf16(uint32_t k) {
  return (k & 0x8000u) ? (uint16_t)(k & 0x7FFFu) : (uint16_t)(~k & 0xFFFFu);
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based -----------------------------------------
__device__ __host__ __forceinline__ void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
  uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __host__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// Exp(1) draw for (seed, step, head, element) as the bf16 value `probs / q` divides by
// (inference.py:24-27).  One Philox block per element: counter = (element, head, step, 0).
__device__ __forceinline__ float exp1_noise(uint64_t seed, uint32_t step, uint32_t head, uint32_t elem) {
  uint32_t c[4] = {elem, head, step, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  float u = ((float)(c[0] >> 9) + 0.5f) * (1.0f / 8388608.0f);   // 23 bits + 1/2: strictly inside (0,1) and exact in fp32, so q > 0
  return rbf(-logf(u));
}

// ---- optional in-kernel timeline (DUALAR_TIMELINE=1): block 0 / thread 0 of every kernel of the step graph
// stamps %globaltimer and clock64 at entry, after the dependency wait, after the prologue and at exit
struct Timeline { unsigned long long *buf; int slot; };   // buf[slot*8 + {0..3}] = globaltimer ns, {4..7} = clock64
__device__ __forceinline__ void tl_stamp(const Timeline &t, int k) {
  if (t.buf && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
    t.buf[t.slot * 8 + k] = g; t.buf[t.slot * 8 + 4 + k] = (unsigned long long)clock64();
  }
}

}  // namespace da

// ---- device-resident request state: everything a graph replay needs lives here ------------------
struct DAState {
  int pos;          // position of the input column (index into KV cache / seq)
  int n_gen;        // token columns produced so far in this request
  int max_gen;      // stop after this many
  int prompt_len;
  int done;         // EOS seen or limit reached -> further replays are no-ops
  int use_penalty;  // repetition penalty on for this step (off for the prefill-produced token)
  int err;          // device-side fault flag
  int loop_mode;    // 1: token loop owned by the engine (finish_step advances the state)
  int cpu_sem;      // scalar-operand semantics of the reference's CPU path instead of its CUDA path (see dualar_set_option)
  float temperature, top_p, rep_penalty;
  unsigned int step_ctr;           // Philox step counter
  unsigned int phase_ctr;          // running phase counter of the persistent kernels: source of the unit tags
  unsigned long long seed;
  const bf16 *noise;               // explicit noise for this step or nullptr
  long long noise_stride;          // elements per step (loop mode advances `noise`)
  int tok_in[DA_MAX_ROWS];         // input column of this step
  int tok_out[DA_MAX_ROWS];        // result of this step
  int win[DA_MAX_ROWS * DA_WIN];   // repetition window (C+1, 16) (inference.py:186-191)
  int nucleus[DA_MAX_ROWS];        // kept-set size per head (diagnostic)
  // scratch tickets / counters, all reset by their last user
  unsigned int attn_ticket[DA_MAX_KV_HEADS];
  unsigned int head_ticket;
  unsigned int sel_ticket;
  unsigned int n_cand;
  unsigned int fast_ticket;
  unsigned long long s_fix;        // slow head: sum of exp(z - max) over the vocabulary, 2^-40 fixed point (order-free)
};
