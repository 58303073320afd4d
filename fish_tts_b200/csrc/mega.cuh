// mega.cuh -- the whole dual-AR decode step (decode_one_token_ar, inference.py:83-155) as ONE persistent kernel.
//
// One CTA per SM (cooperative launch => co-resident), 16 compute warps + 1 producer warp.  A decode step is a static
// table of ~346 dependent PHASES (embedding -> 28 x {wqkv, split-KV attention, merge, wo, w1/w3, w2} -> LM head in 16
// parts interleaved with fast pass 0 -> slow sampler -> 9 fast passes x 4 layers x {wqkv, attention+wo, w1/w3, w2} + 9
// fast heads with their samplers; the first wqkv of passes >= 1 comes from a per-code table, misc_kernels.cuh).
// Three mechanisms keep HBM / L2 busy while the dependency chain advances:
//
//   * WEIGHT RING.  What a CTA will read from global memory is known before the step starts: its contiguous slice
//     of every weight matrix and (from st->pos) its K/V tiles.  The producer warp walks that list and streams it into a
//     shared-memory ring with the bulk-copy engine (cp.async.bulk + mbarrier complete_tx, the 1-D TMA path), running as
//     far ahead of the compute warps as the ring allows (entries are released by mbarrier arrivals).  The weight stream
//     therefore never waits for a hand-over; a phase's dot products read shared memory only.
//   * TAGGED UNITS.  Activations cross SMs as 32-bit units = bf16 value | 16-bit phase tag (64-bit units = fp32 value |
//     32-bit tag for the attention partials and sampler statistics), written with ONE store; the consumer polls until
//     the tag is the producing phase's.  Data and "ready" arrive together: no fence, no atomic, no grid barrier between
//     phases (NCCL's LL protocol, on-chip through L2).
//   * every CTA polls the COMPLETE input of every GEMV phase, so passing a phase proves that all CTAs finished the
//     previous one; buffers are reused without further synchronisation.
//
// Numerics: a row's dot product is summed in the CANONICAL ORDER of mma_chunk() below (128-element chunks, two chains of
// four bf16 MMAs each, chunk partials added in chunk order); every other formula is the one of the per-phase kernels
// (attention.cuh, sampler.cuh, misc_kernels.cuh), whose FMA chains sum in a different fp32 order: the two paths agree to
// bf16 rounding (tests/test_gpu_parity.py::test_mega_kernel_vs_per_phase_kernels), and both samplers are exact on the
// logits they are given.  All spins are bounded: a lost hand-over raises the device fault flag instead of hanging the GPU.
#pragma once
#include "attention.cuh"
#include "common.cuh"
#include "gemv.cuh"
#include "misc_kernels.cuh"
#include "sampler.cuh"
#include <type_traits>

namespace da {

#define DA_M_CWARPS 16               // 16 compute warps + the producer warp.  The 17th warp puts five warps on one scheduler, which caps
#define DA_M_CTHREADS 512            // ptxas at 96 registers (a few spilled kernel-scope values); 15 + 1 warps (128 registers, no spills)
#define DA_M_THREADS 544             // measured 2.5% slower: 16 heads / 16 units per phase then take two rounds.  Everything below is
                                     // written against these macros, so either configuration builds.
#define DA_M_PPW ((DA_TILE + DA_M_CWARPS - 1) / DA_M_CWARPS)   // positions of a 64-row K/V tile per warp
#define DA_M_NB 32                  // ring entries in flight (mbarrier pairs)
#define DA_M_ENTRY_BYTES 16384      // target size of one ring entry of a GEMV phase
#define DA_M_MAX_PHASES 400
#define DA_M_MAXL 48                // slow layers
#define DA_M_MAXFL 8                // fast layers
#define DA_SPIN_LIMIT (1 << 18)      // bounded spins: a lost hand-over raises the fault flag after ~0.1 s instead of hanging the GPU
#define DA_M_REP 1                  // replicas of every broadcast unit vector (CTA b polls replica b % DA_M_REP).  Measured
                                    // (tests/cuda/handover_bench2.cu): ONE copy is fastest -- 0.80 us per all-to-all hand-over of
                                    // 1024 units at 148 CTAs vs 1.03 us with 4 copies; the extra stores cost more than the
                                    // spread of the pollers saves

enum { MK_GEMV = 0, MK_ATTN = 1, MK_MERGE = 2, MK_HSTAT = 3, MK_HCAND = 4, MK_PREFILL_END = 5 };
enum { MP_PLAIN = 0, MP_RMSNORM = 1, MP_FASTATTN = 2, MP_EMBED = 3 };   // MP_EMBED: token embedding computed in place, then RMSNorm
enum { ME_STORE = 0, ME_RESIDUAL = 1, ME_SWIGLU = 2, ME_SLOWLOGITS = 3, ME_FASTLOGITS = 4 };
enum { MF_KEEP = 1, MF_SAVE0 = 2, MF_SAVE1 = 4, MF_RES0 = 8, MF_RES1 = 16, MF_T0 = 32 };   // MF_T0: q | k | v of this attention come from the code table

struct MPhase {
  const bf16 *W, *bias, *norm_w;
  const uint32_t *in;    // input units
  uint32_t *out;         // output units
  int rows, K;
  short in_ph;           // phase that wrote `in` (its tag)
  unsigned char kind, pro, epi, layer, pos, flags;
  short pq, prem;        // row pairs per CTA: floor and remainder of (rows / 2) / grid (the first `prem` CTAs take one more)
  unsigned char part, nparts, cgroup, pad2_;   // nparts > 1: this entry handles only the part-th slice of the CTA's tiles (LM head, see build_mega)
                                               // cgroup: consecutive 128-element chunks one warp computes per unit (1, or more when a CTA
                                               // would otherwise hold more than 16 units and half the warps would go round twice)
};

struct MegaSmem { uint32_t bars, chg, xb, xbh, raw, scratch, work, part, pcnt, lg, kvs, ring, total; };
struct MegaArgs {
  int n_phases;
  // slow attention (llama.py:242-282)
  const bf16 *rope; int nh, nkv, hd, S, nsplit_max; float eps, sf;
  bf16 *kc[DA_M_MAXL], *vc[DA_M_MAXL]; const bf16 *qn[DA_M_MAXL], *kn[DA_M_MAXL];
  unsigned long long *part_o, *part_ml;     // 64-bit units: [nkv][nsplit_max][G][hd], [nkv][nsplit_max][G][2]
  // embedding (llama.py:409-429)
  const bf16 *emb, *cb_emb; int dim, vocab, codebook_size, num_codebooks, sem_begin, sem_end, scale_cb; float inv_sqrt, sqrt_c;
  // slow head + sampler
  bf16 *logits, *logits_raw; unsigned long long *hmax, *hcs, *cand; float delta; int n_rows_tok, head_pq, head_prem;
  // fast stack
  const bf16 *frope; const bf16 *fqn[DA_M_MAXFL], *fkn[DA_M_MAXFL]; int fl, fnh, fnkv, fhd, ncb; float fscale;
  const bf16 *fast_emb; const bf16 *t0; int fdim, fv; uint32_t *u_fin; bf16 *flogits_raw, *flogits; long long noise_off0;
  bf16 *fkv;             // fast K/V rows of the current token, one private copy per CTA: [cta][layer][pos][k | v][nkv * hd] (L2-resident)
  // end of step
  int *seq; int seq_stride, im_end_id;
  DAState *st; unsigned long long *tl; int tl_slots;
  unsigned long long *tl2;   // optional per-CTA stamps [phase][cta][staged, done] (DUALAR_TIMELINE=1): who is the slowest CTA of a phase?
  // test hook (dualar_debug_sample): when set, the logits epilogues take the linear's output from these buffers instead of the dot
  // product, so that penalty, statistics, candidate list and samplers of THIS kernel run on caller-supplied logits
  const bf16 *force_slow, *force_fast;      // [vocab] / [(num_codebooks - 1) * fv], or null
  unsigned int *phase_ctr;   // running phase counter = source of the unit tags; NEVER reset (a request must not see the previous one's tags)
  // shared-memory plan
  int kmax, lg_rows, work_bytes, kv_bytes, ring_bytes;
  MegaSmem plan;         // shared-memory offsets: read from parameter space at every use -- thirteen 64-bit pointers held live across
                         // the phase loop cost ~25 registers under a 96-register cap and made ptxas spill in the hot loops
  MPhase table[DA_M_MAX_PHASES];
};

// ---- units --------------------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ unsigned long long ld_poll8(const unsigned long long *p) {
  unsigned long long r;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_unit(uint32_t *p, uint32_t u) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(u) : "memory"); }
__device__ __forceinline__ void st_unit8(void *p, unsigned long long u) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(u) : "memory"); }
__device__ __forceinline__ unsigned long long make_unit8(uint32_t payload, uint32_t tag) { return ((unsigned long long)payload << 32) | tag; }
__device__ __forceinline__ bool tags_ok(const uint4 &u, uint32_t tag) {
  return ((u.x & 0xFFFFu) == tag) & ((u.y & 0xFFFFu) == tag) & ((u.z & 0xFFFFu) == tag) & ((u.w & 0xFFFFu) == tag);
}
// poll 8 consecutive units (one 8-element chunk) until every tag matches; false on timeout.  No sleep between attempts: a
// polling warp is stalled on the load, not issuing, and any back-off only delays the discovery (handover_bench.cu)
__device__ __noinline__ bool poll_chunk(const uint32_t *p, uint32_t tag, float *f) {
  uint4 a, b;
  int it = 0;
  for (;;) {
    a = ld_poll4(p); b = ld_poll4(p + 4);
    if (tags_ok(a, tag) & tags_ok(b, tag)) break;
    if (++it >= DA_SPIN_LIMIT) break;
  }
  f[0] = unit_val(a.x); f[1] = unit_val(a.y); f[2] = unit_val(a.z); f[3] = unit_val(a.w);
  f[4] = unit_val(b.x); f[5] = unit_val(b.y); f[6] = unit_val(b.z); f[7] = unit_val(b.w);
  return it < DA_SPIN_LIMIT;
}
// two 4-unit groups per thread, `lo` and `hi` half a vector apart: both loads of a warp are fully coalesced (16 sectors per
// request instead of 32 half-used ones), which is worth ~0.15 us per hand-over at 148 pollers (handover_bench2.cu)
__device__ int g_poll_ns = 0;      // back-off between poll attempts (experiments: DUALAR_POLL_NS)
__device__ int g_l2_window = 0;      // 1: the fast stack is covered by a cudaAccessPolicyWindow on the kernel node (engine.cu capture());
                                     // its bulk copies then carry no per-copy cache hint so that the window's property applies
__device__ int g_keep_mode = 0;      // what the other (1 - keep_frac) lines of the fast stack ask for: 0 evict_first, 1 evict_unchanged (DUALAR_KEEP_MODE)
__device__ float g_keep_frac = 0.7f;  // fraction of the fast-stack lines that ask L2 for evict_last (DUALAR_KEEP_FRAC).  The stack
                                      // (109 MB) does not fit next to the slow stream: asking for all of it thrashes (ncu: 2.2 GB
                                      // DRAM reads per token, L2 hit 34%); 0.7 keeps a stable subset (1.69 GB, 46%)
__device__ __forceinline__ bool poll_pair(const uint32_t *lo, const uint32_t *hi, uint32_t tag, float *f) {
  uint4 a, b;
  int it = 0;
  const int ns = g_poll_ns;
  for (;;) {
    a = ld_poll4(lo); b = ld_poll4(hi);
    if (tags_ok(a, tag) & tags_ok(b, tag)) break;
    if (++it >= DA_SPIN_LIMIT) break;
    if (ns) __nanosleep(ns);
  }
  f[0] = unit_val(a.x); f[1] = unit_val(a.y); f[2] = unit_val(a.z); f[3] = unit_val(a.w);
  f[4] = unit_val(b.x); f[5] = unit_val(b.y); f[6] = unit_val(b.z); f[7] = unit_val(b.w);
  return it < DA_SPIN_LIMIT;
}
// poll one 64-bit unit; payload in the high word
__device__ __noinline__ bool poll8(const unsigned long long *p, uint32_t tag, uint32_t &payload) {
  unsigned long long v;
  int it = 0;
  for (;;) {
    v = ld_poll8(p);
    if ((uint32_t)v == tag) break;
    if (++it >= DA_SPIN_LIMIT) break;
    __nanosleep(20);
  }
  payload = (uint32_t)(v >> 32);
  return it < DA_SPIN_LIMIT;
}
// three 64-bit units whose loads are issued together: one L2 round trip once all of them are there
__device__ __noinline__ bool poll8x3(const unsigned long long *p0, const unsigned long long *p1, const unsigned long long *p2, uint32_t tag,
                                     uint32_t &v0, uint32_t &v1, uint32_t &v2) {
  unsigned long long a, b, c;
  int it = 0;
  for (;;) {
    a = ld_poll8(p0); b = ld_poll8(p1); c = ld_poll8(p2);
    if (((uint32_t)a == tag) & ((uint32_t)b == tag) & ((uint32_t)c == tag)) break;
    if (++it >= DA_SPIN_LIMIT) break;
  }
  v0 = (uint32_t)(a >> 32); v1 = (uint32_t)(b >> 32); v2 = (uint32_t)(c >> 32);
  return it < DA_SPIN_LIMIT;
}
// bounded mbarrier wait that does not burn issue slots: a hot try_wait loop by every waiting warp was ~25% of all
// instructions executed by the kernel (ncu) and slowed the warps that had work on the same scheduler
__device__ __forceinline__ bool mbar_wait_idle(uint64_t *bar, uint32_t parity, unsigned ns) {
  uint32_t done = 0;
  for (int it = 0; it < DA_SPIN_LIMIT; ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(ns);
  }
  return done != 0;
}
// shared-memory counter with acquire-release semantics at CTA scope: orders the partial sums written before it (by the
// whole warp, through __syncwarp) against the reads of whoever sees the final count -- without two full MEMBAR.SC
__device__ __forceinline__ int atom_add_acq_rel_cta(volatile int *p, int v) {
  int old;
  asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(const_cast<int *>(p))), "r"(v) : "memory");
  return old;
}
typedef BlockNamed<1, DA_M_CTHREADS> CBlock;           // the 512 compute threads
__device__ __forceinline__ void cbar() { CBlock::sync(); }
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

#define DA_G_IPT 8   // fast heads: warps 0..3 of CTA 0 hold the <= 1024 logits, 8 per thread (named barrier 2)

// ---- the static read list: what CTA `bid` streams through its ring in phase `ph` ------------------------------------------
// Producer and compute warps evaluate the same functions, so they agree on the entry sequence without communicating.
// A GEMV phase deals contiguous row ranges to the CTAs (whole w1/w3 pairs); a CTA walks its range in TILES of 16 rows, one
// ring entry per tile.  Each row lands by its own bulk copy at a stride of 2K + 16 bytes, so that the eight row addresses
// of an ldmatrix 8x8 block fall into eight different 16-byte bank groups (conflict-free without swizzling).
#define DA_M_PT 4                   // partial-sum slots (tiles whose chunk partials may be outstanding at once)
struct GemvPart { int r0, nr, nt; };   // first row, rows of this CTA, tiles
__device__ __forceinline__ GemvPart gemv_part(int rows, int pq, int prem, int bid) {
  GemvPart g;
  const int p0 = bid * pq + min(bid, prem), p1 = p0 + pq + (bid < prem ? 1 : 0);
  g.r0 = 2 * p0; g.nr = min(rows, 2 * p1) - g.r0; if (g.nr < 0) g.nr = 0;
  g.nt = (g.nr + 15) >> 4;
  return g;
}
__device__ __forceinline__ uint32_t row_stride(int K) { return 2u * (uint32_t)K + 16u; }

// CANONICAL ORDER of a row's dot product in this kernel family: K is cut into chunks of 128 elements; a chunk partial is
// the sum of two fp32 accumulators, each a chain of four m16n8k16 bf16 tensor-core MMAs starting from zero (the even and
// the odd 16-element steps of the chunk, k ascending; the x vector is replicated into all eight B columns); the row value
// is the sum of the chunk partials in chunk order.  Which warp computes which chunk, and which other rows share the tile,
// does not change the result.  All fragment loads are issued before the first MMA and the two chains are independent, so
// a unit costs about four dependent MMA latencies instead of eight load + MMA round trips.
__device__ __forceinline__ void mma_chunk(uint32_t tile_addr, uint32_t RS, int nrows, const uint32_t *xw, int chunk, int lane, float &v_lo, float &v_hi) {
  int row = lane & 15; if (row >= nrows) row = 0;
  const uint32_t addr = tile_addr + (uint32_t)row * RS + (uint32_t)(((lane >> 4) << 3) + (chunk << 7)) * 2u;
  const uint32_t *xp = xw + (chunk << 6) + (lane & 3);
  float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll
  for (int hb = 0; hb < 8; hb += 4) {        // fragments of four steps at a time (24 registers; all eight spilled under the 96-register cap)
    uint32_t af[4][4], bfr[4][2];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(af[s][0]), "=r"(af[s][1]), "=r"(af[s][2]), "=r"(af[s][3]) : "r"(addr + (hb + s) * 32));
      bfr[s][0] = xp[(hb + s) * 8]; bfr[s][1] = xp[(hb + s) * 8 + 4];
    }
#pragma unroll
    for (int s = 0; s < 4; s += 2) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(e0), "+f"(e1), "+f"(e2), "+f"(e3) : "r"(af[s][0]), "r"(af[s][1]), "r"(af[s][2]), "r"(af[s][3]), "r"(bfr[s][0]), "r"(bfr[s][1]));
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(o0), "+f"(o1), "+f"(o2), "+f"(o3) : "r"(af[s + 1][0]), "r"(af[s + 1][1]), "r"(af[s + 1][2]), "r"(af[s + 1][3]), "r"(bfr[s + 1][0]), "r"(bfr[s + 1][1]));
    }
  }
  v_lo = e0 + o0; v_hi = e2 + o2;     // rows lane/4 and lane/4 + 8 (every B column holds x, so every lane of a group has them)
}
__device__ __forceinline__ void unpack4(const uint2 &u, float *f) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
}
// four consecutive activation values -> packed bf16, one 8-byte store at element offset e
__device__ __forceinline__ void store4_xb(bf16 *xb, int e, const float *f) {
  uint2 u;
  u.x = (uint32_t)f2bits(f[0]) | ((uint32_t)f2bits(f[1]) << 16); u.y = (uint32_t)f2bits(f[2]) | ((uint32_t)f2bits(f[3]) << 16);
  *reinterpret_cast<uint2 *>(xb + e) = u;
}
// dpl (1, 2 or 4) consecutive bf16 values as floats with ONE load; the rest of f[4] is zero
__device__ __forceinline__ void ld_bf16_dpl(const bf16 *p, int dpl, float *f) {
  f[0] = f[1] = f[2] = f[3] = 0.f;
  if (dpl == 4) unpack4(*reinterpret_cast<const uint2 *>(p), f);
  else if (dpl == 2) { const uint32_t u = *reinterpret_cast<const uint32_t *>(p); f[0] = __uint_as_float(u << 16); f[1] = __uint_as_float(u & 0xffff0000u); }
  else f[0] = bf2f(*p);
}
// eight consecutive activation values (bf16-exact floats) -> packed bf16, one 16-byte store
__device__ __forceinline__ void store_chunk_xb(bf16 *xb, int c, const float *f) {
  uint4 u;
  u.x = (uint32_t)f2bits(f[0]) | ((uint32_t)f2bits(f[1]) << 16); u.y = (uint32_t)f2bits(f[2]) | ((uint32_t)f2bits(f[3]) << 16);
  u.z = (uint32_t)f2bits(f[4]) | ((uint32_t)f2bits(f[5]) << 16); u.w = (uint32_t)f2bits(f[6]) | ((uint32_t)f2bits(f[7]) << 16);
  reinterpret_cast<uint4 *>(xb)[c] = u;
}
struct AttnPart { int active, g, split, t0, t1, nsplit_eff, L; };
__device__ __forceinline__ AttnPart attn_part(int pos, int nkv, int nsplit_max, int bid) {
  AttnPart p;
  p.L = pos + 1;
  const int n_tiles = (p.L + DA_TILE - 1) / DA_TILE;
  const int nsplit = min(nsplit_max, n_tiles);
  const int tps = (n_tiles + nsplit - 1) / nsplit;
  p.nsplit_eff = (n_tiles + tps - 1) / tps;
  p.g = bid % nkv; p.split = bid / nkv;
  p.active = (bid < nkv * nsplit_max) && (p.split < p.nsplit_eff);
  p.t0 = p.split * tps; p.t1 = min(n_tiles, p.t0 + tps);
  return p;
}
// rows of tile t that come from the cache (everything but the position being written now)
__device__ __forceinline__ int tile_old_rows(int t, int pos) {
  const int r0 = t * DA_TILE, r1 = min(pos + 1, r0 + DA_TILE);
  return min(r1, pos) - r0;
}

struct RingCursor { uint32_t off, seq; };
// place an entry of `size` bytes; entries never wrap (the tail of the ring is skipped and charged to the entry)
__device__ __forceinline__ uint32_t ring_place(RingCursor &r, uint32_t size, uint32_t ring_bytes, uint32_t &charged) {
  uint32_t pad = 0;
  if (r.off + size > ring_bytes) { pad = ring_bytes - r.off; r.off = 0; }
  const uint32_t at = r.off;
  r.off += size; charged = size + pad; r.seq += 1;
  return at;
}

// ---- shared-memory plan (host and device) -----------------------------------------------------------------------------------
static inline __host__ __device__ MegaSmem mega_smem_plan(int kmax, int dim_max, int lg_rows, int work_bytes, int kv_bytes, int ring_bytes) {
  MegaSmem m; uint32_t o = 0;
  m.bars = o; o += 2 * DA_M_NB * 8;
  m.chg = o; o += DA_M_NB * 4;
  m.xb = o; o += 2u * (uint32_t)kmax * 2;                                  // two staging buffers, by phase parity
  m.xbh = o; o += (uint32_t)dim_max * 2;                                   // staging buffer of the LM head: survives the interleaved fast phases
  m.raw = o; o += 2u * (uint32_t)dim_max * 4;
  m.scratch = o; o += 160 * 4;
  m.work = o; o += ((uint32_t)work_bytes + 15u) & ~15u;
  m.part = o; o += (uint32_t)DA_M_PT * (uint32_t)(kmax >> 7) * 16 * 4;    // [slot][chunk][row]
  m.pcnt = o; o += 2 * DA_M_PT * 4;                                        // chunk counters, fold generations
  m.lg = o; o += (((uint32_t)lg_rows * 2) + 15u) & ~15u;
  m.kvs = o; o += ((uint32_t)kv_bytes + 15u) & ~15u;
  o = (o + 127u) & ~127u;
  m.ring = o; o += (uint32_t)ring_bytes;
  m.total = o;
  return m;
}

__device__ __forceinline__ void tl_mark(const MegaArgs &a, int slot, int k) {
  if (a.tl && blockIdx.x == 0 && threadIdx.x == 0 && slot < a.tl_slots) {
    unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
    a.tl[slot * 8 + k] = g;
  }
}

__device__ __forceinline__ unsigned long long gtime() { unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); return g; }
// unconditional variant for the instrumented thread of CTA 0 (producer stamps, accumulated waits)
__device__ __forceinline__ void tl_put(const MegaArgs &a, int slot, int k, unsigned long long v) {
  if (a.tl && blockIdx.x == 0 && slot < a.tl_slots) a.tl[slot * 8 + k] = v;
}

// block sum over the first `nwarps` compute warps (the ones that hold the vector), named barrier 3 over just those warps: the
// other warps are not held up and the barrier is cheaper.  All participants get the result; scratch >= 16 floats, alternate
// by parity.  Fixed order.
__device__ __forceinline__ float cblock_sum(float v, float *scratch, int lane, int w, int nwarps) {
  v = warp_sum(v);
  if (lane == 0) scratch[w] = v;
  named_bar(3, nwarps * 32);
  float t = 0.f;
  for (int i = 0; i < nwarps; ++i) t += scratch[i];
  return t;
}

// =====================================================================================================================
// TL: timeline instrumentation compiled in (DUALAR_TIMELINE=1 runs only).  FULL: the decode step (LM head, samplers, fast AR
// loop); !FULL = one prefill position (slow stack only).  The variants exist because code the kernel never executes still
// slowed it down: the instrumentation alone cost 3%, five unused attention variants 10% (instruction-cache footprint).
template <bool TL, bool FULL>
__global__ void __launch_bounds__(DA_M_THREADS, 1) mega_kernel(const __grid_constant__ MegaArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  DAState *st = a.st;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int bid = blockIdx.x, grid = gridDim.x;
  const int dim_max = a.dim > a.fdim ? a.dim : a.fdim;
#define sm_full (reinterpret_cast<uint64_t *>(sm + a.plan.bars))
#define sm_empty (reinterpret_cast<uint64_t *>(sm + a.plan.bars) + DA_M_NB)
#define sm_chg (reinterpret_cast<uint32_t *>(sm + a.plan.chg))
#define sm_xb2 (reinterpret_cast<bf16 *>(sm + a.plan.xb))
#define sm_xbh (reinterpret_cast<bf16 *>(sm + a.plan.xbh))
#define sm_raw (reinterpret_cast<float *>(sm + a.plan.raw))
#define sm_scratch (reinterpret_cast<float *>(sm + a.plan.scratch))
#define sm_work (sm + a.plan.work)
#define sm_part (reinterpret_cast<float *>(sm + a.plan.part))
#define sm_pcnt (reinterpret_cast<volatile int *>(sm + a.plan.pcnt))
#define sm_pgen (reinterpret_cast<volatile int *>(sm + a.plan.pcnt) + DA_M_PT)
#define sm_lg (reinterpret_cast<uint16_t *>(sm + a.plan.lg))
#define sm_kvs (reinterpret_cast<bf16 *>(sm + a.plan.kvs))
#define sm_ring (sm + a.plan.ring)
  const uint32_t ring_bytes = (uint32_t)a.ring_bytes;

  if (TL) tl_mark(a, 0, 0);
  if (tid < 2 * DA_M_PT) sm_pcnt[tid] = 0;      // chunk counters and fold generations of the partial-sum slots
  if (tid == 0) {
    for (int i = 0; i < DA_M_NB; ++i) { mbar_init(&sm_full[i], 1); mbar_init(&sm_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // request state: read once, before anything of this step can have changed it
  const int done = *reinterpret_cast<const volatile int *>(&st->done);
  const int pos = *reinterpret_cast<const volatile int *>(&st->pos);
  const uint32_t tag_base = *reinterpret_cast<const volatile unsigned int *>(a.phase_ctr);
  __syncthreads();
  if (done) return;
  const int nph = a.n_phases;
  const int G = a.nh / a.nkv;

  // =================================================== producer ===========================================================
  if (w == DA_M_CWARPS) {
    // the fast stack is re-read num_codebooks times per token: ask L2 to keep (a fraction of) it, stream everything else
    uint64_t pol_keep;
    { const float fr = g_keep_frac;
      if (g_keep_mode == 1) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_unchanged.b64 %0, %1;" : "=l"(pol_keep) : "f"(fr));
      else asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol_keep) : "f"(fr)); }
    const uint64_t pol_stream = policy_evict_first();
    const bool use_window = g_l2_window != 0;
    RingCursor rc = {0u, 0u};
    uint32_t used = 0, tail = 0;
    bool ok = true;
    // lane 0 owns the sm_ring bookkeeping; returns the smem offset of the new entry (rc.seq - 1 is its index)
    auto reserve = [&](uint32_t size) -> uint32_t {
      RingCursor probe = rc; uint32_t charged;
      ring_place(probe, size, ring_bytes, charged);
      while (used + charged > ring_bytes || rc.seq - tail >= DA_M_NB) {
        ok = mbar_wait_idle(&sm_empty[tail % DA_M_NB], (tail / DA_M_NB) & 1u, 100) && ok;
        used -= sm_chg[tail % DA_M_NB]; ++tail;
      }
      const uint32_t at = ring_place(rc, size, ring_bytes, charged);
      sm_chg[(rc.seq - 1) % DA_M_NB] = charged; used += charged;
      return at;
    };
    for (int ph = 0; ph < nph; ++ph) {
      const MPhase &d = a.table[ph];
      if (TL && a.tl && lane == 0) tl_put(a, 1 + ph, 4, gtime());
      if (d.kind == MK_GEMV) {
        const GemvPart gp = gemv_part(d.rows, d.pq, d.prem, bid);
        const uint64_t pol = (d.flags & MF_KEEP) ? pol_keep : pol_stream;
        const uint32_t RS = row_stride(d.K), row_bytes = 2u * (uint32_t)d.K;
        const int t_lo = d.nparts > 1 ? gp.nt * d.part / d.nparts : 0, t_hi = d.nparts > 1 ? gp.nt * (d.part + 1) / d.nparts : gp.nt;
        for (int t = t_lo; t < t_hi; ++t) {
          const int n = min(16, gp.nr - 16 * t);
          uint32_t at = 0, bi = 0;
          if (lane == 0) {
            at = reserve((uint32_t)n * RS); bi = (rc.seq - 1) % DA_M_NB;
            mbar_expect_tx(&sm_full[bi], (uint32_t)n * row_bytes);
          }
          at = __shfl_sync(0xffffffffu, at, 0); bi = __shfl_sync(0xffffffffu, bi, 0);
          if (lane < n) {
            if (use_window && (d.flags & MF_KEEP)) bulk_g2s_nohint(sm_ring + at + (uint32_t)lane * RS, d.W + (size_t)(gp.r0 + 16 * t + lane) * d.K, row_bytes, &sm_full[bi]);
            else bulk_g2s(sm_ring + at + (uint32_t)lane * RS, d.W + (size_t)(gp.r0 + 16 * t + lane) * d.K, row_bytes, &sm_full[bi], pol);
          }
        }
      } else if (d.kind == MK_ATTN) {
        const AttnPart ap = attn_part(pos, a.nkv, a.nsplit_max, bid);
        if (ap.active && lane == 0) {
          for (int t = ap.t0; t < ap.t1; ++t) {
            const int n_old = tile_old_rows(t, pos);
            if (n_old <= 0) continue;
            const uint32_t bytes = (uint32_t)n_old * a.hd * 2u;
            const uint32_t at = reserve(2 * bytes);
            uint64_t *fb = &sm_full[(rc.seq - 1) % DA_M_NB];
            mbar_expect_tx(fb, 2 * bytes);
            const size_t src = ((size_t)ap.g * a.S + (size_t)t * DA_TILE) * a.hd;
            bulk_g2s(sm_ring + at, a.kc[d.layer] + src, bytes, fb, pol_stream);
            bulk_g2s(sm_ring + at + bytes, a.vc[d.layer] + src, bytes, fb, pol_stream);
          }
        }
        __syncwarp();
      }
      if (TL && a.tl && lane == 0) tl_put(a, 1 + ph, 5, gtime());
    }
    if (!ok) st->err = 2;
    return;
  }

  // =================================================== compute warps =======================================================
  auto tag_of = [&](int ph) { return (uint32_t)((tag_base + (uint32_t)ph) % 65535u) + 1u; };
  auto tag32_of = [&](int ph) { return tag_base + (uint32_t)ph + 1u; };
  RingCursor rc = {0u, 0u};
  bool ok = true;
  unsigned long long wait_ns = 0ull;
  int sparity = 0;
  // (request parameters -- repetition penalty, temperature, top-p, noise source -- are read from the state inside the logits /
  //  sampler phases that use them: values held live across the whole phase loop are what ptxas spills first)
  // the next sm_ring entry: every warp advances the cursor; only warps that read the entry wait for it to land.  An entry is
  // released by ONE arrival -- lane 31 of the warp that folds the tile (all its units have been read by then; lane 31 never
  // has global stores in flight, which a release-type operation would wait for), or one thread after the CTA barrier that
  // ends an attention tile.
  auto place = [&](uint32_t size, uint32_t &bar_idx, uint32_t &parity) -> uint32_t {
    uint32_t charged;
    const uint32_t at = ring_place(rc, size, ring_bytes, charged);
    const uint32_t i = rc.seq - 1;
    bar_idx = i % DA_M_NB; parity = (i / DA_M_NB) & 1u;
    return at;
  };
  long long wait_cyc = 0;      // cycles this thread spent waiting for sm_ring entries (reported per CTA when the timeline is on)
  auto landed = [&](uint32_t bar_idx, uint32_t parity) {
    if (TL && a.tl) { const long long c0 = clock64(); ok = mbar_wait_idle(&sm_full[bar_idx], parity, 20) && ok; const long long dc = clock64() - c0; wait_cyc += dc; wait_ns += (unsigned long long)dc / 2; }
    else ok = mbar_wait_idle(&sm_full[bar_idx], parity, 20) && ok;
  };
  // folds completed on each partial-sum slot in earlier phases (the generation counters in shared memory never reset)
  int gen_base[DA_M_PT];
#pragma unroll
  for (int i = 0; i < DA_M_PT; ++i) gen_base[i] = 0;
  // softmax statistics of the slow head carried from MK_GEMV(ME_SLOWLOGITS) to MK_HSTAT / MK_HCAND
  float h_m = 0.f, h_wmax = -INFINITY; int h_cnt = 0;

  for (int ph = 0; ph < nph; ++ph) {
    const MPhase &d = a.table[ph];
    const uint32_t tag = tag_of(ph), in_tag = tag_of(d.in_ph);
    const uint32_t *in = d.in;
    if (TL) tl_mark(a, 1 + ph, 0);
    if (TL && a.tl && tid == 0 && ph > 0) tl_put(a, ph, 6, wait_ns);
    wait_ns = 0ull;
    float *sc = sm_scratch + sparity * 16; sparity ^= 1;

    if (d.kind == MK_GEMV) {
      const int K = d.K, nchunk = K >> 7;
      const GemvPart gp = gemv_part(d.rows, d.pq, d.prem, bid);
      // staging buffer of this phase; the other one may still be read by a slow warp of the previous phase.  A phase that comes
      // in parts (the LM head, interleaved with the fast pass 0) stages once, into its own buffer.
      const bool parted = d.nparts > 1;
      bf16 *xb = parted ? sm_xbh : sm_xb2 + (size_t)(ph & 1) * a.kmax;
      const int t_lo = parted ? gp.nt * d.part / d.nparts : 0, t_hi = parted ? gp.nt * (d.part + 1) / d.nparts : gp.nt;
      if (!parted || d.part == 0) {
      // ---- (A) stage the input vector as packed bf16 (every activation is a bf16 value) ------------------------------------------
      if (FULL && d.pro == MP_FASTATTN) {
        // fast-layer attention for position d.pos (llama.py:246-251, 285-309), recomputed by every CTA.  After the RoPE
        // barrier one WARP owns one head end to end (scores, softmax, P@V), so nothing else synchronises the CTA.
        const int nh = a.fnh, nkv = a.fnkv, hd = a.fhd, qd = nh * hd, kd = nkv * hd, FG = nh / nkv;
        const int p = d.pos, P = p + 1;
        float *q = reinterpret_cast<float *>(sm_work), *kcur = q + qd, *vcur = kcur + kd;
        // The K / V rows of this token's earlier codebook positions live in a per-CTA sm_scratch in global memory (L2): keeping
        // all layers' rows in shared memory cost 80 KB of the weight sm_ring.  This layer's rows are requested BEFORE the poll
        // for the new q | k | v, so the L2 round trip hides behind the hand-over wait.
        bf16 *kv_l = sm_kvs;                                             // [pos][k | v][kd] bf16, this layer only
        bf16 *kv_g = a.fkv + ((size_t)bid * a.fl + d.layer) * a.ncb * 2 * kd;
        uint4 pre[3];
        const int n16 = p * 2 * kd / 8;                                // 16-byte pieces of the rows of positions < p
#pragma unroll
        for (int i = 0; i < 3; ++i) { const int c = tid + i * DA_M_CTHREADS; if (c < n16) pre[i] = __ldcg(reinterpret_cast<const uint4 *>(kv_g) + c); }
        // Thread c polls elements [4c, 4c+4) and [N/2 + 4c, +4) of q | k | v.  A head vector is hd/4 consecutive lanes of one warp
        // in either half, so qk-norm and RoPE run in registers on the polled values; q goes to shared memory as fp32, the new
        // K / V row straight into the caches as bf16 -- one CTA barrier before the scores.
        {
          const int c = tid, N = qd + 2 * kd, LH = hd >> 2;
          const bf16 *rope_row = a.frope + (size_t)p * hd;
          const bool any_norm = a.fqn[d.layer] || a.fkn[d.layer];
          if ((c & ~31) * 8 < N) {                            // warp-uniform: the shuffles below need whole warps
            const bool act = c * 8 < N;
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = 0.f;
            if (d.flags & MF_T0) {
              // first layer of a pass >= 1: the input is the embedding of the code just drawn, so q | k | v come from the table
              // (one tagged unit carries the code; every thread polls it, then reads its 8 values and, for the residual of
              //  this layer's wo, its share of the embedding row)
              uint32_t u = 0u;
              { int it = 0; for (;;) { u = ld_poll1(in); if ((u & 0xFFFFu) == in_tag) break; if (++it >= DA_SPIN_LIMIT) { ok = false; break; } } }
              uint32_t code = u >> 16; if (code >= (uint32_t)a.codebook_size) code = 0u;
              if (act) {
                const bf16 *row = a.t0 + (size_t)code * N;
                unpack4(*reinterpret_cast<const uint2 *>(row + c * 4), t); unpack4(*reinterpret_cast<const uint2 *>(row + N / 2 + c * 4), t + 4);
              }
              if (c * 8 < a.fdim) {
                float xr[8];
                const bf16 *er = a.fast_emb + (size_t)code * a.fdim + c * 8;
                unpack4(*reinterpret_cast<const uint2 *>(er), xr); unpack4(*reinterpret_cast<const uint2 *>(er + 4), xr + 4);
                *reinterpret_cast<float4 *>(sm_raw + c * 8) = make_float4(xr[0], xr[1], xr[2], xr[3]);
                *reinterpret_cast<float4 *>(sm_raw + c * 8 + 4) = make_float4(xr[4], xr[5], xr[6], xr[7]);
              }
            } else if (act) ok = poll_pair(in + c * 4, in + N / 2 + c * 4, in_tag, t) && ok;
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              float *tt = t + 4 * hlf;
              const int e = hlf * (N / 2) + c * 4;            // global element; region and head follow from it
              const bool is_q = e < qd, is_k = !is_q && e < qd + kd;
              const int eh = e & (hd - 1);                    // element inside the head vector (qd, kd are multiples of hd)
              const bf16 *nw = is_q ? a.fqn[d.layer] : (is_k ? a.fkn[d.layer] : nullptr);
              if (any_norm) {
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) ss = fmaf(tt[j], tt[j], ss);
                for (int o = LH >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                if (act && nw) {
                  const float rstd = rsqrtf(ss * (1.0f / (float)hd) + a.eps);
                  float gw[4]; unpack4(*reinterpret_cast<const uint2 *>(nw + eh), gw);
#pragma unroll
                  for (int j = 0; j < 4; ++j) tt[j] = rbf(__fmul_rn(__fmul_rn(tt[j], rstd), gw[j]));
                }
              }
              if (act && (is_q || is_k)) {                    // RoPE, interleaved pairs (llama.py:606-618); no fma contraction
                float rr[4]; unpack4(*reinterpret_cast<const uint2 *>(rope_row + eh), rr);
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                  const float x0 = tt[j], x1 = tt[j + 1];
                  tt[j] = rbf(__fsub_rn(__fmul_rn(x0, rr[j]), __fmul_rn(x1, rr[j + 1])));
                  tt[j + 1] = rbf(__fadd_rn(__fmul_rn(x1, rr[j]), __fmul_rn(x0, rr[j + 1])));
                }
              }
              if (act) {
                if (is_q) *reinterpret_cast<float4 *>(q + e) = make_float4(tt[0], tt[1], tt[2], tt[3]);
                else {                                        // this position's K / V row: shared-memory cache + per-CTA scratch
                  const size_t off = ((size_t)p * 2 + (is_k ? 0 : 1)) * kd + (size_t)(e - qd - (is_k ? 0 : kd));
                  const uint2 pk = make_uint2((uint32_t)f2bits(tt[0]) | ((uint32_t)f2bits(tt[1]) << 16), (uint32_t)f2bits(tt[2]) | ((uint32_t)f2bits(tt[3]) << 16));
                  *reinterpret_cast<uint2 *>(kv_l + off) = pk; *reinterpret_cast<uint2 *>(kv_g + off) = pk;
                }
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) { const int c = tid + i * DA_M_CTHREADS; if (c < n16) reinterpret_cast<uint4 *>(kv_l)[c] = pre[i]; }
        if (TL && a.tl && tid == 0) tl_put(a, 512 + ph, 0, gtime());
        cbar();
        if (TL && a.tl && tid == 0) tl_put(a, 512 + ph, 1, gtime());
        if (TL && a.tl && tid == 0) tl_put(a, 512 + ph, 2, gtime());
        // one HALF-warp per head (30 half-warps for <= 16 heads: every head in one round), both halves run the same code
        float *prs = vcur + kd;                                     // [nh][16] scores, then probabilities
        const int sl = lane & 15, hsel = lane & 16;
        for (int hb = 2 * w; hb < ((nh + 1) & ~1); hb += 2 * DA_M_CWARPS) {
          const int h_raw = hb + (lane >> 4);
          const bool hlive = h_raw < nh;
          const int h = hlive ? h_raw : nh - 1, g = h / FG;
          const float *qq = q + h * hd;
          // scores: 4 lanes per position, hd/4 elements each; bf16(q @ k^T), then bf16(* scale)   (llama.py:304)
          const int prt = sl & 3, ne = hd >> 2;
          for (int j0 = 0; j0 < P; j0 += 4) {
            const int j = j0 + (sl >> 2);
            float acc = 0.f;
            if (j < P) {
              const bf16 *kk = kv_l + ((size_t)j * 2 + 0) * kd + g * hd + prt * ne;
              const float *qp = qq + prt * ne;
#pragma unroll 4
              for (int x = 0; x < ne; x += 2) {
                const float2 kf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(kk + x));
                acc = fmaf(qp[x], kf.x, acc); acc = fmaf(qp[x + 1], kf.y, acc);
              }
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (prt == 0 && j < P) prs[h * 16 + j] = rbf(__fmul_rn(rbf(acc), a.fscale));
          }
          __syncwarp();
          // softmax over j <= pos in fp32, rounded to bf16 (llama.py:305-306; masked columns are exp(-inf) = 0): lane sl <-> position sl
          const float sc_j = sl < P ? prs[h * 16 + sl] : -INFINITY;
          float m = sc_j;
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          const float ej = sl < P ? expf(sc_j - m) : 0.f;
          float sum = ej;
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const float pj = sl < P ? rbf(ej / sum) : 0.f;
          // y = bf16(p @ v)   (llama.py:309): lane sl owns dims sl*dph .. +dph of its head (dph = hd / 16: 2, 4 or 8)
          const int dph = hd >> 4;
          for (int d2 = 0; d2 < dph; d2 += 2) {
            const int d0 = sl * dph + d2;
            float y0 = 0.f, y1 = 0.f;
            for (int jj = 0; jj < P; ++jj) {
              const float pv = __shfl_sync(0xffffffffu, pj, hsel | jj);
              const float2 vf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(kv_l + ((size_t)jj * 2 + 1) * kd + g * hd + d0));
              y0 = fmaf(pv, vf.x, y0); y1 = fmaf(pv, vf.y, y1);
            }
            if (hlive) *reinterpret_cast<uint32_t *>(xb + h * hd + d0) = (uint32_t)f2bits(y0) | ((uint32_t)f2bits(y1) << 16);
          }
        }
        if (TL && a.tl && tid == 0) tl_put(a, 512 + ph, 3, gtime());
      } else {
        // thread c < K/8 owns elements [4c, 4c+4) and [K/2 + 4c, K/2 + 4c + 4)
        const int c = tid, e_lo = 4 * c, e_hi = (K >> 1) + 4 * c;
        const bool mine = c * 8 < K;
        float v[8], g[8];
        const bool normed = d.pro == MP_RMSNORM || d.pro == MP_EMBED;
        if (mine && normed) { unpack4(*reinterpret_cast<const uint2 *>(d.norm_w + e_lo), g); unpack4(*reinterpret_cast<const uint2 *>(d.norm_w + e_hi), g + 4); }
        if (d.pro == MP_EMBED) {
          // token + codebook embedding (llama.py:409-429), every CTA computes the whole vector
          if (mine) {
            int tok = st->tok_in[0];
            if (tok < 0 || tok >= a.vocab) { tok = 0; st->err = 1; }
            const bool is_sem = tok >= a.sem_begin && tok <= a.sem_end;
            float te[8];
            unpack4(*reinterpret_cast<const uint2 *>(a.emb + (size_t)tok * a.dim + e_lo), te); unpack4(*reinterpret_cast<const uint2 *>(a.emb + (size_t)tok * a.dim + e_hi), te + 4);
            float vq[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) vq[j] = 0.f;
            if (is_sem) {
              for (int i = 0; i < a.num_codebooks; ++i) {       // stack(...).sum(dim=1): fp32 accumulate, one rounding
                int cc = st->tok_in[i + 1];
                if (cc < 0 || cc >= a.codebook_size) { cc = 0; st->err = 1; }
                const bf16 *rowp = a.cb_emb + ((size_t)cc + (size_t)i * a.codebook_size) * a.dim;
                float ce[8]; unpack4(*reinterpret_cast<const uint2 *>(rowp + e_lo), ce); unpack4(*reinterpret_cast<const uint2 *>(rowp + e_hi), ce + 4);
#pragma unroll
                for (int j = 0; j < 8; ++j) vq[j] += ce[j];
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) vq[j] = rbf(vq[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x = rbf(te[j] + vq[j]);
              if (a.scale_cb && is_sem) x = st->cpu_sem ? rbf(__fdiv_rn(x, a.sqrt_c)) : rbf(__fmul_rn(x, a.inv_sqrt));
              v[j] = x;
            }
          }
        } else if (mine) ok = poll_pair(in + e_lo, in + e_hi, in_tag, v) && ok;
        if (mine && (d.flags & (MF_SAVE0 | MF_SAVE1))) {
          float *dst = sm_raw + ((d.flags & MF_SAVE1) ? dim_max : 0);
          *reinterpret_cast<float4 *>(dst + e_lo) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4 *>(dst + e_hi) = make_float4(v[4], v[5], v[6], v[7]);
        }
        const int nw_vec = (K / 8 + 31) >> 5;      // warps that hold a piece of the vector
        if (normed && w < nw_vec) {
          float ss = 0.f;
          if (mine) {
#pragma unroll
            for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
          }
          ss = cblock_sum(ss, sc, lane, w, nw_vec);
          const float inv = rsqrtf(ss * (1.0f / (float)K) + a.eps);
          if (mine) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = rbf(__fmul_rn(rbf(__fmul_rn(v[j], inv)), g[j]));
          }
        }
        if (mine) { store4_xb(xb, e_lo, v); store4_xb(xb, e_hi, v + 4); }
      }
      cbar();
      }      // staging
      if (TL) tl_mark(a, 1 + ph, 1);
      if (TL && a.tl2 && tid == 0) a.tl2[((size_t)ph * grid + bid) * 2] = gtime();

      // ---- (B) units: (tile, chunk of 128 elements) -> warp (tile * nchunk + chunk) % 16.  A unit leaves 16 partial sums in
      //      shared memory; the warp that completes a tile's last chunk folds them in chunk order and runs the epilogue,
      //      one row per lane -- no CTA-wide barrier between the dot products and the stores.
      int pen_id = -1;      // penalised ids of the logits epilogues live in lanes 0..15 of every warp
      float rp_eff = 1.f; int use_pen = 0;
      if (FULL && (d.epi == ME_FASTLOGITS || d.epi == ME_SLOWLOGITS)) { rp_eff = eff_rep_penalty(st); use_pen = st->use_penalty; }
      if (FULL && d.epi == ME_FASTLOGITS && use_pen && lane < DA_WIN) pen_id = st->win[(d.pos + 1) * DA_WIN + lane];
      if (FULL && d.epi == ME_SLOWLOGITS && use_pen && lane < a.n_rows_tok) pen_id = st->win[lane * DA_WIN];     // previous_tokens[:, 0]
      if (d.epi == ME_SLOWLOGITS && (!parted || d.part == 0)) h_wmax = -INFINITY;      // the CTA maximum runs over all parts of the head
      const float *resv = sm_raw + ((d.flags & MF_RES1) ? dim_max : 0);
      const uint32_t RS = row_stride(K);
      const uint32_t *xw = reinterpret_cast<const uint32_t *>(xb);
      {
        int u = w;
        const int cg = d.cgroup ? d.cgroup : 1, upt = nchunk / cg;      // units per tile
        for (int t = t_lo; t < t_hi; ++t) {
          const int n = min(16, gp.nr - 16 * t), tl = t - t_lo;      // tl: tile index within this phase entry (units, slots, generations)
          uint32_t bi, par;
          const uint32_t at = place((uint32_t)n * RS, bi, par);
          if (u >= (tl + 1) * upt) continue;           // no unit of this warp in the tile
          landed(bi, par);
          if (TL && a.tl && tid == 0 && tl == 0) tl_put(a, 512 + ph, 4, gtime());
          const int slot = tl & (DA_M_PT - 1);
          int gb = gen_base[0];
#pragma unroll
          for (int i = 1; i < DA_M_PT; ++i) if (slot == i) gb = gen_base[i];
          const int gen_need = gb + tl / DA_M_PT;
          float *pslot = sm_part + (size_t)slot * (a.kmax >> 7) * 16;
          for (; u < (tl + 1) * upt; u += DA_M_CWARPS) {
            const int c0 = (u - tl * upt) * cg;
            // the slot is free once the tile DA_M_PT before this one has been folded
            // (checked for every tile: a phase entry without a staging barrier -- the later LM-head parts -- can start while the
            //  previous phase is still folding on this slot)
            { int it = 0; while (sm_pgen[slot] != gen_need) { if (++it >= DA_SPIN_LIMIT) { ok = false; break; } __nanosleep(20); } }
            for (int c = c0; c < c0 + cg; ++c) {      // every chunk keeps its own partial: the fold adds them in chunk order (canonical)
              float v_lo, v_hi;
              mma_chunk(smem_u32(sm_ring + at), RS, n, xw, c, lane, v_lo, v_hi);
              if (TL && a.tl && tid == 0 && tl == 0) tl_put(a, 512 + ph, 5, gtime() + (unsigned long long)(0.f * (v_lo + v_hi)));
              if ((lane & 3) == 0) { pslot[c * 16 + (lane >> 2)] = v_lo; pslot[c * 16 + 8 + (lane >> 2)] = v_hi; }
            }
            __syncwarp();
            int last = 0;
            if (lane == 31) last = (atom_add_acq_rel_cta(&sm_pcnt[slot], cg) == nchunk - cg);      // publishes this warp's partials, observes the others'
            last = __shfl_sync(0xffffffffu, last, 31);
            if (TL && a.tl && tid == 0 && tl == 0) tl_put(a, 512 + ph, 6, gtime());
            if (last) {
              // every unit of the tile is done: fold in chunk order (lane r owns row r), free the slot and the sm_ring entry
              const int r = lane & 15, row = gp.r0 + 16 * t + r;
              const bool live = lane < 16 && r < n;
              float v = 0.f;
              for (int c8 = 0; c8 < nchunk; c8 += 8) {        // K is a multiple of 256: chunks come in groups of 8; loads first, then the ordered adds
                float pv[8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) pv[cc] = (c8 + cc < nchunk) ? pslot[(c8 + cc) * 16 + r] : 0.f;
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) if (c8 + cc < nchunk) v += pv[cc];
              }
              __syncwarp();
              if (lane == 31) { sm_pcnt[slot] = 0; __threadfence_block(); sm_pgen[slot] = gen_need + 1; mbar_arrive(&sm_empty[bi]); }
              if (live && d.bias) v += bf2f(d.bias[row]);
              if (d.epi == ME_STORE) {
                if (live) st_unit(d.out + row, make_unit(v, tag));
              } else if (d.epi == ME_RESIDUAL) {
                if (live) st_unit(d.out + row, make_unit(resv[row] + rbf(v), tag));
              } else if (d.epi == ME_SWIGLU) {
                const float up = __shfl_down_sync(0xffffffffu, v, 1);      // row 2j = w1 (gate), row 2j+1 = w3 (up)
                if (live && !(r & 1)) {
                  const float gg = rbf(v), uu = rbf(up);
                  const float sg = rbf(gg / (1.0f + expf(-gg)));
                  st_unit(d.out + (row >> 1), make_unit(__fmul_rn(sg, uu), tag));
                }
              } else if (FULL) {
                float z = rbf(v);
                if (d.epi == ME_SLOWLOGITS ? a.force_slow != nullptr : a.force_fast != nullptr) {
                  if (live) z = d.epi == ME_SLOWLOGITS ? bf2f(a.force_slow[row]) : bf2f(a.force_fast[(size_t)(d.pos - 1) * a.fv + row]);
                }
                bool hit = false;
#pragma unroll
                for (int i = 0; i < DA_WIN; ++i) hit |= (__shfl_sync(0xffffffffu, pen_id, i) == row);
                if (d.epi == ME_FASTLOGITS) {
                  const size_t lo = (size_t)(d.pos - 1) * a.fv;
                  if (live) {
                    a.flogits_raw[lo + row] = f2bf(z);
                    if (hit) z = penalise(z, rp_eff);
                    a.flogits[lo + row] = f2bf(z); st_unit(d.out + row, make_unit(z, tag));
                  }
                } else if (live) {   // ME_SLOWLOGITS: logits stay in this CTA's shared memory for the sampler phases; global copies for read-back / fallback
                  a.logits_raw[row] = f2bf(z);
                  if (hit) z = penalise(z, rp_eff);
                  a.logits[row] = f2bf(z); sm_lg[16 * t + r] = f2bits(z); h_wmax = fmaxf(h_wmax, z);
                }
              }
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < DA_M_PT; ++i) gen_base[i] += (t_hi - t_lo + DA_M_PT - 1 - i) / DA_M_PT;
      if (TL) tl_mark(a, 1 + ph, 2);
      if (TL && a.tl2 && tid == 0) { a.tl2[((size_t)ph * grid + bid) * 2 + 1] = gtime(); a.tl2[(size_t)DA_M_MAX_PHASES * 160 * 2 + (size_t)ph * grid + bid] = wait_ns; }

      // ---- (C) heads ------------------------------------------------------------------------------------------------------
      if (FULL && d.epi == ME_SLOWLOGITS && (!parted || d.part == d.nparts - 1)) {
        // CTA max of the penalised logits -> 64-bit unit; the fence makes this CTA's global logits visible to whoever
        // has seen the unit (needed by the whole-vocabulary fallback sampler only)
        const float wmax = warp_max(h_wmax);
        if (lane == 0) sc[w] = wmax;
        __threadfence();
        cbar();
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < DA_M_CWARPS; ++i) m = fmaxf(m, sc[i]);
        if (tid == 0) st_unit8(a.hmax + bid, make_unit8(__float_as_uint(m), tag32_of(ph)));
        h_cnt = gp.nr;
      }
      if (FULL && d.epi == ME_FASTLOGITS && bid == 0) {
        // CTA 0 draws the code with all 16 warps (two logits per thread, binned nucleus sampler) and publishes its embedding
        // as the next pass's input
        unsigned long long *scr = reinterpret_cast<unsigned long long *>(sm_work);
        const int V = a.fv;
        uint32_t it2[2];
        Red r = {0ull, 0, -1};
        {
          const bool mine = tid * 2 < V;     // V is a multiple of 8 (codebook sizes are)
          uint32_t u0 = 0u, u1 = 0u;
          if (mine) {
            int it = 0;
            for (;;) {
              asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(u0), "=r"(u1) : "l"(d.out + tid * 2) : "memory");
              if (((u0 & 0xFFFFu) == tag) & ((u1 & 0xFFFFu) == tag)) break;
              if (++it >= DA_SPIN_LIMIT) { ok = false; break; }
            }
          }
          const uint32_t k0 = bf16_key((uint16_t)(u0 >> 16)), k1 = bf16_key((uint16_t)(u1 >> 16));
          it2[0] = mine ? (((0xFFFFu - k0) << 16) | (uint32_t)(tid * 2)) : 0xFFFFFFFFu;
          it2[1] = mine ? (((0xFFFFu - k1) << 16) | (uint32_t)(tid * 2 + 1)) : 0xFFFFFFFFu;
          if (mine) r.m = (int)max(k0, k1);
        }
        SampleParams spm;
        int par = 0;
        if (TL && a.tl && tid == 0) tl_put(a, 1024 + ph, 0, gtime());
        r = block_reduce<CBlock>(r, scr, par);
        spm.m = bits2f(key_bf16((uint32_t)r.m));
        Red es = {0ull, 0, -1};      // sum of exp terms as 2^-40 fixed point: order-free, identical to the per-phase path
#pragma unroll
        for (int i = 0; i < 2; ++i) if (it2[i] != 0xFFFFFFFFu) es.s += (unsigned long long)(expf(bits2f(key_bf16(0xFFFFu - (it2[i] >> 16))) - spm.m) * DA_FIX2_SCALE);
        es = block_reduce<CBlock>(es, scr, par);
        spm.S = __ull2float_rn(es.s) * (1.0f / DA_FIX2_SCALE);
        spm.T_bf = eff_temperature(st);
        spm.c_max = cmax_from_top_p(st->top_p);
        const NoiseSrc nsrc = noise_src(st);
        cbar();
        if (TL && a.tl && tid == 0) tl_put(a, 1024 + ph, 1, gtime());
        uint32_t tok = sample_binned<2, DA_M_CTHREADS, CBlock>(it2, (uint32_t)V, true, nullptr, spm, nsrc, (uint32_t)d.pos, a.noise_off0 + (long long)(d.pos - 1) * a.fv,
                                                                &st->nucleus[d.pos], reinterpret_cast<uint32_t *>(scr + 256), scr);
        if (tok >= (uint32_t)a.codebook_size) { tok = a.codebook_size - 1; if (tid == 0) st->err = 3; }
        if (tid == 0) st->tok_out[d.pos + 1] = (int)tok;
        if (TL && a.tl && tid == 0) tl_put(a, 1024 + ph, 2, gtime());
        if (d.pos < a.ncb - 1) {
          if (a.t0) { if (tid == 0) st_unit(a.u_fin, (tok << 16) | tag); }      // the code itself: the next pass looks its first q | k | v up
          else for (int dd = tid; dd < a.fdim; dd += DA_M_CTHREADS) st_unit(a.u_fin + dd, make_unit(bf2f(a.fast_emb[(size_t)tok * a.fdim + dd]), tag));
        }
      }
      if (TL && a.tl && tid == 0) tl_put(a, 1024 + ph, 3, gtime());
      // No CTA barrier ends a plain GEMV phase: the staging buffer alternates, the partial-sum slots are handed over by their
      // generation counters, and the first barrier of the next phase's staging cannot be passed before every fold of this
      // phase is done.  The head phases keep one (sm_work / sm_lg are reused by the sampler phases that follow).
      if (FULL && ((d.epi == ME_SLOWLOGITS && (!parted || d.part == d.nparts - 1)) || d.epi == ME_FASTLOGITS)) cbar();

    } else if (d.kind == MK_ATTN) {
      // ---- slow-layer attention for one query position: split-KV flash-decode (llama.py:242-282 under SDPBackend.MATH) ---------
      // CTA (kv head g, split) walks its 64-position tiles from the sm_ring.  Inside the CTA the positions of a tile are dealt to
      // the 16 warps; every warp keeps its own running (max, sum, output) in registers -- lane l owns dims hd/32*l.. of all G
      // query heads -- and the warps meet once, in shared memory, after the last tile.  fp32 throughout, like torch's math
      // SDPA (q and k both scaled by sqrt(scale)); bf16 only at the very end (merge phase).
      const AttnPart ap = attn_part(pos, a.nkv, a.nsplit_max, bid);
      if (ap.active) {
        const int hd = a.hd, g = ap.g, L = ap.L, dpl = hd >> 5;       // dims per lane (1, 2 or 4)
        const int qd = a.nh * hd, kd = a.nkv * hd;
        float *q = reinterpret_cast<float *>(sm_work);
        float *knew = q + G * hd, *vnew = knew + hd;
        float *pm = vnew + hd, *pl = pm + DA_M_CWARPS * G, *po = pl + DA_M_CWARPS * G;       // per-warp partials: [16][G], [16][G], [16][G*hd]
        bf16 *knb = reinterpret_cast<bf16 *>(po + (size_t)DA_M_CWARPS * G * hd), *vnb = knb + hd;   // the new K / V row as bf16, laid out like a tile row
        const bool owns_new = (pos / DA_TILE) >= ap.t0 && (pos / DA_TILE) < ap.t1;
        {   // one 8-unit chunk per thread: G*hd/8 chunks of q, then hd/8 of the new k and of the new v.  A head vector is hd/8
            // consecutive lanes of one warp, so qk-norm (llama.py:246-251) and RoPE (:606-618) run in registers on the polled
            // values: no shared-memory round trip and a single CTA barrier before the tile walk.
          const int LH = hd >> 3, nq = G * LH, nk = owns_new ? LH : 0;
          const int c = tid;
          if ((c & ~31) < nq + 2 * nk) {                      // warp-uniform: the shuffles below need whole warps
            const bool act = c < nq + 2 * nk, is_q = c < nq, is_k = act && !is_q && c < nq + nk;
            const int e0 = (c & (LH - 1)) * 8;                // first element of this chunk inside its head vector
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = 0.f;
            if (act) {
              const uint32_t *src = is_q ? in + (size_t)g * G * hd + (size_t)c * 8
                                  : is_k ? in + qd + (size_t)g * hd + (size_t)(c - nq) * 8 : in + qd + kd + (size_t)g * hd + (size_t)(c - nq - nk) * 8;
              ok = poll_chunk(src, in_tag, t) && ok;
            }
            const bf16 *nw = is_q ? a.qn[d.layer] : (is_k ? a.kn[d.layer] : nullptr);
            if (a.qn[d.layer] || a.kn[d.layer]) {             // nn.RMSNorm over the head vector: one rounding, after the weight multiply
              float ss = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) ss = fmaf(t[j], t[j], ss);
              for (int o = LH >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
              if (nw) {
                const float rstd = rsqrtf(ss * (1.0f / (float)hd) + a.eps);
                float gw[8]; unpack4(*reinterpret_cast<const uint2 *>(nw + e0), gw); unpack4(*reinterpret_cast<const uint2 *>(nw + e0 + 4), gw + 4);
#pragma unroll
                for (int j = 0; j < 8; ++j) t[j] = rbf(__fmul_rn(__fmul_rn(t[j], rstd), gw[j]));
              }
            }
            if (is_q || is_k) {                               // separate mul / mul / add kernels in the reference: no fma contraction
              float rr[8]; const bf16 *rope_row = a.rope + (size_t)pos * hd + e0;
              unpack4(*reinterpret_cast<const uint2 *>(rope_row), rr); unpack4(*reinterpret_cast<const uint2 *>(rope_row + 4), rr + 4);
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                const float x0 = t[j], x1 = t[j + 1];
                t[j] = rbf(__fsub_rn(__fmul_rn(x0, rr[j]), __fmul_rn(x1, rr[j + 1])));
                t[j + 1] = rbf(__fadd_rn(__fmul_rn(x1, rr[j]), __fmul_rn(x0, rr[j + 1])));
              }
            }
            if (is_q) {
#pragma unroll
              for (int j = 0; j < 8; ++j) q[c * 8 + j] = __fmul_rn(t[j], a.sf);      // q * sqrt(scale), fp32
            } else if (act) {                                 // KVCache.update (llama.py:142-149): both rows are bf16 values already
              uint4 pk;
              pk.x = (uint32_t)f2bits(t[0]) | ((uint32_t)f2bits(t[1]) << 16); pk.y = (uint32_t)f2bits(t[2]) | ((uint32_t)f2bits(t[3]) << 16);
              pk.z = (uint32_t)f2bits(t[4]) | ((uint32_t)f2bits(t[5]) << 16); pk.w = (uint32_t)f2bits(t[6]) | ((uint32_t)f2bits(t[7]) << 16);
              bf16 *cache = is_k ? a.kc[d.layer] : a.vc[d.layer];
              *reinterpret_cast<uint4 *>((is_k ? knb : vnb) + e0) = pk;
              *reinterpret_cast<uint4 *>(cache + ((size_t)g * a.S + pos) * hd + e0) = pk;
            }
          }
        }
        if (TL && a.tl && tid == 0) tl_put(a, 512 + ph, 0, gtime());
        cbar();
        if (TL) tl_mark(a, 1 + ph, 1);
        if (TL && a.tl && tid == 0) tl_put(a, 512 + ph, 1, gtime());
        // Running (max, sum, output) of every warp live in its private slice of shared memory (pm / pl / po) and are pulled into
        // registers for two query heads at a time, so one code path serves every GQA group size and head dimension (dims per lane
        // dpl = hd / 32 <= 4) without per-shape instantiations -- unexecuted variants measurably slowed the whole kernel down.
        // (the first tile starts from the empty state in registers; only a CTA without tiles has to write it)
        if (ap.t1 <= ap.t0) {
          for (int h = 0; h < G; ++h) {
            if (lane == 0) { pm[w * G + h] = -INFINITY; pl[w * G + h] = 0.f; }
            for (int i = 0; i < dpl; ++i) po[((size_t)w * G + h) * hd + lane * dpl + i] = 0.f;
          }
          __syncwarp();
        }
        for (int t = ap.t0; t < ap.t1; ++t) {
          const int r0 = t * DA_TILE, r1 = min(L, r0 + DA_TILE), nrow = r1 - r0;
          const int n_old = min(r1, pos) - r0;
          uint32_t bi = 0; const bf16 *kt = nullptr, *vt = nullptr;
          if (n_old > 0) {
            const uint32_t bytes = (uint32_t)n_old * hd * 2u;
            uint32_t par;
            const uint32_t at = place(2 * bytes, bi, par);
            landed(bi, par);
            kt = reinterpret_cast<const bf16 *>(sm_ring + at); vt = reinterpret_cast<const bf16 *>(sm_ring + at + bytes);
          }
          if (TL && a.tl && tid == 0) tl_put(a, 512 + ph, 3, gtime());
          for (int h0 = 0; h0 < G; h0 += 2) {
            float qr[2][4], m_run[2], l_run[2], o_acc[2][4];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const bool hv = h0 + hh < G;
              const int h = hv ? h0 + hh : h0;
              const bool first = t == ap.t0;
              m_run[hh] = first ? -INFINITY : pm[w * G + h]; l_run[hh] = first ? 0.f : pl[w * G + h];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const bool iv = hv && i < dpl;
                qr[hh][i] = iv ? q[h * hd + lane * dpl + i] : 0.f;
                o_acc[hh][i] = (iv && !first) ? po[((size_t)w * G + h) * hd + lane * dpl + i] : 0.f;
              }
            }
            // A warp owns at most four positions of a 64-row tile (rows w, w+16, w+32, w+48).  Scores: 8 lanes per position, each
            // lane a strided set of 4-element pieces of the head vector for both query heads -- a 3-step reduction per score instead
            // of a 32-lane one, and one exp per lane.  The probabilities are then handed to all lanes (4 shuffles per head) for the
            // P@V accumulation, where a lane owns dpl dims of the output as before.
            static_assert(DA_M_PPW == 4, "the score layout deals four position slots to the 32 lanes");
            const int psl = lane >> 3, dl = lane & 7;
            const int jme = w + psl * DA_M_CWARPS;
            float s2[2] = {0.f, 0.f};
            {
              const bf16 *krow = (jme < n_old ? kt + (size_t)jme * hd : knb);      // (rows past nrow read the new-row buffer; masked below)
              const float *q0 = q + (size_t)h0 * hd, *q1 = q + (size_t)(h0 + 1 < G ? h0 + 1 : h0) * hd;
              for (int e = dl * 4; e < hd; e += 32) {
                float kf[4]; unpack4(*reinterpret_cast<const uint2 *>(krow + e), kf);
                const float4 qa = *reinterpret_cast<const float4 *>(q0 + e), qb = *reinterpret_cast<const float4 *>(q1 + e);
#pragma unroll
                for (int i = 0; i < 4; ++i) kf[i] = __fmul_rn(kf[i], a.sf);
                s2[0] = fmaf(qa.x, kf[0], s2[0]); s2[0] = fmaf(qa.y, kf[1], s2[0]); s2[0] = fmaf(qa.z, kf[2], s2[0]); s2[0] = fmaf(qa.w, kf[3], s2[0]);
                s2[1] = fmaf(qb.x, kf[0], s2[1]); s2[1] = fmaf(qb.y, kf[1], s2[1]); s2[1] = fmaf(qb.z, kf[2], s2[1]); s2[1] = fmaf(qb.w, kf[3], s2[1]);
              }
            }
            float pj[DA_M_PPW][2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              float sv = s2[hh];
              sv += __shfl_xor_sync(0xffffffffu, sv, 1); sv += __shfl_xor_sync(0xffffffffu, sv, 2); sv += __shfl_xor_sync(0xffffffffu, sv, 4);
              if (!(jme < nrow && h0 + hh < G)) sv = -INFINITY;
              float mx = fmaxf(sv, __shfl_xor_sync(0xffffffffu, sv, 8));
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
              for (int jj = 0; jj < DA_M_PPW; ++jj) pj[jj][hh] = __shfl_sync(0xffffffffu, pme, jj * 8);
            }
#pragma unroll
            for (int jj = 0; jj < DA_M_PPW; ++jj) {
              const int j = w + jj * DA_M_CWARPS;
              if (j < nrow) {
                float vf[4];
                ld_bf16_dpl((j < n_old ? vt + j * hd : vnb) + lane * dpl, dpl, vf);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  l_run[hh] += pj[jj][hh];
#pragma unroll
                  for (int i = 0; i < 4; ++i) o_acc[hh][i] = fmaf(pj[jj][hh], vf[i], o_acc[hh][i]);
                }
              }
            }
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
          if (TL && a.tl && tid == 0) tl_put(a, 512 + ph, 4, gtime());
          if (n_old > 0) {      // every warp is past its last read of the tile: one arrival releases the sm_ring entry
            cbar();
            if (tid == DA_M_CTHREADS - 1) mbar_arrive(&sm_empty[bi]);
          }
        }
        if (TL && a.tl && tid == 0) tl_put(a, 512 + ph, 2, gtime());
        // thread e = (h, d) folds the 16 partials in warp order
        cbar();
        const uint32_t t32 = tag32_of(ph);
        unsigned long long *pog = a.part_o + (((size_t)g * a.nsplit_max + ap.split) * G) * hd;
        unsigned long long *pmlg = a.part_ml + (((size_t)g * a.nsplit_max + ap.split) * G) * 2;
        // thread e = (h, d): maximum over the 16 warps, then the 16 scaled partials in warp order (every thread recomputes the
        // scale factors of its head: 16 exps instead of two more CTA barriers)
        for (int e = tid; e < G * hd; e += DA_M_CTHREADS) {
          const int h = e / hd, dd = e - h * hd;
          float m = -INFINITY;
#pragma unroll
          for (int ww = 0; ww < DA_M_CWARPS; ++ww) m = fmaxf(m, pm[ww * G + h]);
          float l = 0.f, o = 0.f;
#pragma unroll
          for (int ww = 0; ww < DA_M_CWARPS; ++ww) {
            const float v = pm[ww * G + h];
            const float sc_w = v == -INFINITY ? 0.f : expf(v - m);
            l = fmaf(pl[ww * G + h], sc_w, l);
            o = fmaf(po[((size_t)ww * G + h) * hd + dd], sc_w, o);
          }
          st_unit8(pog + e, make_unit8(__float_as_uint(o), t32));
          if (dd == 0) { st_unit8(pmlg + h * 2, make_unit8(__float_as_uint(m), t32)); st_unit8(pmlg + h * 2 + 1, make_unit8(__float_as_uint(l), t32)); }
        }
        cbar();
      }
      if (TL) tl_mark(a, 1 + ph, 2);

    } else if (d.kind == MK_MERGE) {
      // ---- merge the split-KV partials in split order (deterministic) and publish y as bf16 units ------------------------------
      const AttnPart ap = attn_part(pos, a.nkv, a.nsplit_max, bid);
      const int hd = a.hd, E = a.nh * hd, ns = ap.nsplit_eff;
      const int e0 = (int)(((long long)E * bid) / grid), e1 = (int)(((long long)E * (bid + 1)) / grid), ne = e1 - e0;
      float *mo = reinterpret_cast<float *>(sm_work), *mm = mo + ne * ns, *ml = mm + ne * ns;    // [ne][ns] each
      const uint32_t t32 = tag32_of(d.in_ph);
      for (int t = tid; t < ne * ns; t += DA_M_CTHREADS) {
        const int i = t / ns, s = t - i * ns, e = e0 + i;
        const int h = e / hd, dd = e - h * hd, g = h / G, hl = h - g * G;
        const size_t base = ((size_t)g * a.nsplit_max + s) * G + hl;
        uint32_t vo, vm, vl;
        ok = poll8x3(a.part_o + base * hd + dd, a.part_ml + base * 2, a.part_ml + base * 2 + 1, t32, vo, vm, vl) && ok;
        mo[t] = __uint_as_float(vo); mm[t] = __uint_as_float(vm); ml[t] = __uint_as_float(vl);
      }
      cbar();
      if (TL) tl_mark(a, 1 + ph, 1);
      if (tid < ne) {
        const int i = tid;
        float m = -INFINITY;
        for (int s = 0; s < ns; ++s) m = fmaxf(m, mm[i * ns + s]);
        float l = 0.f, o = 0.f;
        for (int s = 0; s < ns; ++s) {
          const float sc_s = expf(mm[i * ns + s] - m);
          l = fmaf(ml[i * ns + s], sc_s, l);
          o = fmaf(mo[i * ns + s], sc_s, o);
        }
        st_unit(d.out + e0 + i, make_unit(o / l, tag));
      }
      cbar();
      if (TL) tl_mark(a, 1 + ph, 2);

    } else if (FULL && d.kind == MK_HSTAT) {
      // ---- slow head, stage 2: global max; this CTA's share of S = sum exp(z - m) (2^-40 fixed point) and its candidate count ---
      const uint32_t t32 = tag32_of(d.in_ph);
      float m = -INFINITY;
      if (tid < grid) { uint32_t v; ok = poll8(a.hmax + tid, t32, v) && ok; m = __uint_as_float(v); }
      __threadfence();
      m = warp_max(m);
      if (lane == 0) sc[w] = m;
      cbar();
      m = -INFINITY;
#pragma unroll
      for (int i = 0; i < DA_M_CWARPS; ++i) m = fmaxf(m, sc[i]);
      h_m = m;
      const float thr = m - a.delta;
      Red r = {0ull, 0, -1};
      for (int i = tid; i < h_cnt; i += DA_M_CTHREADS) {
        const float z = bits2f(sm_lg[i]);
        r.s += (unsigned long long)(expf(z - m) * DA_FIX2_SCALE);
        r.c += z >= thr;
      }
      int par = 0;
      r = block_reduce<CBlock>(r, reinterpret_cast<unsigned long long *>(sm_work), par);
      if (tid == 0) {
        const uint32_t o32 = tag32_of(ph);
        st_unit8(a.hcs + 3 * bid + 0, make_unit8((uint32_t)r.c, o32));
        st_unit8(a.hcs + 3 * bid + 1, make_unit8((uint32_t)r.s, o32));
        st_unit8(a.hcs + 3 * bid + 2, make_unit8((uint32_t)(r.s >> 32), o32));
      }
      cbar();
      if (TL) tl_mark(a, 1 + ph, 2);

    } else if (FULL && d.kind == MK_HCAND) {
      // ---- slow head, stage 3: candidates (z >= max - delta) of every CTA into one list; CTA 0 samples (inference.py:47-80) ------
      const uint32_t t32 = tag32_of(d.in_ph), tag30 = tag32_of(ph) & 0x3FFFFFFFu;
      uint32_t *hv = reinterpret_cast<uint32_t *>(sm_work);        // [3*grid] payloads, then 4 words of results
      __shared__ uint32_t s_base, s_n; __shared__ unsigned long long s_S;
      for (int t = tid; t < 3 * grid; t += DA_M_CTHREADS) { uint32_t v; ok = poll8(a.hcs + t, t32, v) && ok; hv[t] = v; }
      cbar();
      if (w == 0) {
        uint32_t base = 0, n = 0; unsigned long long S = 0ull;
        for (int b = lane; b < grid; b += 32) {
          const uint32_t c = hv[3 * b];
          n += c; if (b < bid) base += c;
          S += ((unsigned long long)hv[3 * b + 2] << 32) | hv[3 * b + 1];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { base += __shfl_xor_sync(0xffffffffu, base, o); n += __shfl_xor_sync(0xffffffffu, n, o); S += __shfl_xor_sync(0xffffffffu, S, o); }
        if (lane == 0) { s_base = base; s_n = n; s_S = S; }
      }
      cbar();
      const float m = h_m, thr = m - a.delta;
      const GemvPart gp = gemv_part(a.vocab, a.head_pq, a.head_prem, bid);
      const uint32_t base = s_base, N = s_n;
      if (TL && a.tl && tid == 0) tl_put(a, 1024 + ph, 0, gtime());
      if (N <= DA_CAND_CAP) {
        // this CTA's candidates in vocabulary order: slot numbers increase with the index (ties of the sort need that)
        const int per = (h_cnt + DA_M_CTHREADS - 1) / DA_M_CTHREADS;
        const int i0 = min(h_cnt, tid * per), i1 = min(h_cnt, i0 + per);
        uint32_t cnt = 0;
        for (int i = i0; i < i1; ++i) cnt += bits2f(sm_lg[i]) >= thr;
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
        uint32_t *wtot = hv + 3 * grid + 8;
        if (lane == 31) wtot[w] = inc;
        cbar();
        uint32_t slot = base + inc - cnt;
        for (int i = 0; i < w; ++i) slot += wtot[i];
        for (int i = i0; i < i1; ++i) {
          const uint16_t b = sm_lg[i];
          if (bits2f(b) >= thr) {
            const uint32_t idx = (uint32_t)(gp.r0 + i);
            st_unit8(a.cand + slot, ((unsigned long long)(0xFFFFu - bf16_key(b)) << 48) | ((unsigned long long)idx << 30) | tag30);
            ++slot;
          }
        }
      }
      if (TL && a.tl && tid == 0) tl_put(a, 1024 + ph, 1, gtime());
      if (bid == 0) {
        SampleParams spm;
        spm.m = m; spm.S = __ull2float_rn(s_S) * (1.0f / DA_FIX2_SCALE);
        spm.T_bf = eff_temperature(st);
        spm.c_max = cmax_from_top_p(st->top_p);
        const NoiseSrc nsrc = noise_src(st);
        uint32_t *sm_sort = reinterpret_cast<uint32_t *>(sm_work);
        unsigned long long *scr = reinterpret_cast<unsigned long long *>(sm_work + 16896);
        cbar();   // hv is dead: sm_work becomes the sampler's sm_scratch
        uint32_t idx = 0xFFFFFFFFu;
        // binned sampler over the candidate list (<= 4096 candidates, all 16 warps); wider nuclei take the whole-vocabulary fallback
        if (N >= 1 && N <= 4096) {
          // 8 consecutive entries per thread, fetched with 16-byte polls issued together (entry = inverted key (16) | index (18) | tag (30))
          const unsigned e0 = tid * 8;
          uint32_t it[8];
          int spin = 0;
          for (;;) {
            bool good = true;
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              unsigned long long k0 = 0, k1 = 0;
              if (e0 + i < N) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(k0), "=l"(k1) : "l"(a.cand + e0 + i) : "memory");
              it[i] = (e0 + i < N) ? (((uint32_t)(k0 >> 48) << 16) | (e0 + i)) : 0xFFFFFFFFu;
              it[i + 1] = (e0 + i + 1 < N) ? (((uint32_t)(k1 >> 48) << 16) | (e0 + i + 1)) : 0xFFFFFFFFu;
              if (e0 + i < N) good &= ((uint32_t)(k0 & 0x3FFFFFFFu) == tag30);
              if (e0 + i + 1 < N) good &= ((uint32_t)(k1 & 0x3FFFFFFFu) == tag30);
            }
            if (good || ++spin >= DA_SPIN_LIMIT) break;
          }
          ok = ok && spin < DA_SPIN_LIMIT;
          if (TL && a.tl && tid == 0) tl_put(a, 1024 + ph, 2, gtime());
          idx = sample_binned<8, DA_M_CTHREADS, CBlock>(it, N, (int)N == a.vocab, a.cand, spm, nsrc, 0u, 0ll, &st->nucleus[0], sm_sort, scr);
        }
        cbar();
        if (idx == 0xFFFFFFFFu) {
          __threadfence();
          idx = sample_fallback<CBlock>(a.logits, a.vocab, spm, nsrc, 0u, 0ll, &st->nucleus[0], scr + 192, reinterpret_cast<float *>(scr + 192 + 34));
        }
        // inference.py:123-126: first codebook = semantic id - semantic_begin (clamped at 0); next input = its fast embedding
        if (TL && a.tl && tid == 0) tl_put(a, 1024 + ph, 3, gtime());
        int cb0 = (int)idx - a.sem_begin; if (cb0 < 0) cb0 = 0;
        if (cb0 >= a.codebook_size) { cb0 = a.codebook_size - 1; if (tid == 0) st->err = 3; }
        if (a.t0) { if (tid == 0) st_unit(a.u_fin, ((uint32_t)cb0 << 16) | tag); }
        else for (int dd = tid; dd < a.fdim; dd += DA_M_CTHREADS) st_unit(a.u_fin + dd, make_unit(bf2f(a.fast_emb[(size_t)cb0 * a.fdim + dd]), tag));
        if (tid == 0) { st->tok_out[0] = (int)idx; st->tok_out[1] = cb0; st->n_cand = N; }      // n_cand: diagnostic (candidates of this step)
      }
      cbar();
      if (TL) tl_mark(a, 1 + ph, 2);

    } else if (!FULL && d.kind == MK_PREFILL_END) {
      // one prefill position done: wait for the last layer's output (so every CTA has finished), then load the next prompt column
      if (bid == 0) {
        const int c = tid;
        float v[8];
        if (c * 8 < a.dim) ok = poll_chunk(in + c * 8, in_tag, v) && ok;
        cbar();
      }
    }
  }

  if (TL) tl_mark(a, 0, 3);
  if (TL && a.tl2 && lane == 0) a.tl2[(size_t)DA_M_MAX_PHASES * 148 * 2 + bid * 16 + w] = (unsigned long long)wait_cyc;
  if (!ok && lane == 0) st->err = 4;
  if (bid == 0) {
    cbar();
    if (tid == 0) {
      *a.phase_ctr = tag_base + (unsigned)nph;
      if (!FULL) {
        const int np = pos + 1;
        st->pos = np;
        for (int r = 0; r < a.n_rows_tok; ++r) st->tok_in[r] = a.seq[(size_t)r * a.seq_stride + np];
      } else {
        GemvArgs g; g.st = st; g.seq = a.seq; g.seq_stride = a.seq_stride; g.im_end_id = a.im_end_id; g.n_rows_tok = a.n_rows_tok; g.park_on_done = 0;
        finish_step(g);
      }
    }
  }
}


#undef sm_full
#undef sm_empty
#undef sm_chg
#undef sm_xb2
#undef sm_xbh
#undef sm_raw
#undef sm_scratch
#undef sm_work
#undef sm_part
#undef sm_pcnt
#undef sm_pgen
#undef sm_lg
#undef sm_kvs
#undef sm_ring

}  // namespace da
