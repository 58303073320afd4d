// engine.cu -- host side of libdualar.so: weight arena, KV caches, CUDA-graph step, the C-ABI.
// See include/dualar.h for the contract of every entry point and the reference interface it replaces.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/dualar.h"
#include "attention.cuh"
#include "common.cuh"
#include "gemv.cuh"
#include "misc_kernels.cuh"
#include "mega.cuh"
#include "gemm_tc.cuh"
#include "batch.cuh"
#include "bstep.cuh"

using namespace da;

struct dualar_tc;      // tensor-core path: tensor maps, split-K workspace, prefill columns (batch_host.cuh)
struct dualar_batch;   // batched decode: request slots (batch_host.cuh)

static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
  return code;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(DUALAR_ECUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

struct LayerW {
  bf16 *wqkv = nullptr, *bqkv = nullptr, *wo = nullptr, *bo = nullptr, *qn = nullptr, *kn = nullptr;
  bf16 *w13 = nullptr, *w2 = nullptr, *ffn_norm = nullptr, *attn_norm = nullptr;
  bf16 *kc = nullptr, *vc = nullptr; bool kv_external = false;
};
struct Slot { bf16 *dst; int64_t n; int rows, cols, pitch_rows; bool loaded; };  // pitch_rows: 2 => interleaved halves

struct dualar_engine {
  dualar_config c;
  int device = 0, sms = 0;
  bool finalized = false, request_open = false;
  char *arena = nullptr; size_t arena_bytes = 0, fast_off = 0, fast_bytes = 0;
  std::map<std::string, Slot> slots;
  std::vector<LayerW> slow, fast;
  bf16 *emb = nullptr, *cb_emb = nullptr, *norm = nullptr, *out_w = nullptr, *fast_emb = nullptr, *fast_norm = nullptr, *fast_out = nullptr;
  bf16 *fpi_w = nullptr, *fpi_b = nullptr, *fpi = nullptr; uint32_t *u_fpi = nullptr;      // fast_project_in (fast_dim != dim): weights, projected hidden state
  bf16 *rope = nullptr, *fast_rope = nullptr; bool rope_loaded = false, fast_rope_loaded = false;
  // activations
  bf16 *x = nullptr, *h = nullptr, *qkv = nullptr, *y = nullptr, *act = nullptr, *logits = nullptr, *logits_raw = nullptr;
  bf16 *fin = nullptr, *fbuf[2] = {nullptr, nullptr}, *fh = nullptr, *fqkv = nullptr, *fact = nullptr, *flogits = nullptr, *flogits_raw = nullptr;
  float *part_o = nullptr, *part_ml = nullptr; float2 *partials = nullptr; unsigned long long *cand = nullptr;
  DAState *st = nullptr; int *seq = nullptr; int *h_seq = nullptr; DAState *h_st = nullptr;
  int *h_err_sticky = nullptr, *d_err_sticky = nullptr;   // step mode: fault flag of finished steps, in mapped pinned memory
  void *kv_arena = nullptr;
  int nsplit = 1, fv = 0, gemv_grid = 0;
  float delta = 6.0f; int cpu_sem = 0;
  uint64_t seed = 0; const bf16 *noise = nullptr; int64_t noise_steps = 0;
  cudaStream_t cap_stream = nullptr;
  cudaGraphExec_t g_step = nullptr, g_prefill = nullptr;
  int launches_step = 0, launches_prefill = 0;
  int prompt_len = 0, max_gen = 0;
  std::vector<void *> owned;
  unsigned long long *tl = nullptr; int tl_slots = 0; unsigned long long *tl2 = nullptr;
  // persistent whole-step kernel (mega.cuh): unit buffers, the two phase tables (decode step / one prefill position)
  uint32_t *u_x = nullptr, *u_qkv = nullptr, *u_y = nullptr, *u_h = nullptr, *u_act = nullptr;
  uint32_t *u_fqkv = nullptr, *u_fh = nullptr, *u_fact = nullptr, *u_fx0 = nullptr, *u_fx1 = nullptr, *u_fin = nullptr, *u_flogits = nullptr;
  bf16 *dbg_fast = nullptr;                    // dualar_debug_sample: forced fast-head logits
  bf16 *t0 = nullptr; bool use_t0 = true;      // first-layer q | k | v of the fast stack per code (passes >= 1)
  unsigned long long *m_part_o = nullptr, *m_part_ml = nullptr, *m_hmax = nullptr, *m_hcs = nullptr, *m_cand = nullptr;
  MegaArgs *ma_step = nullptr, *ma_prefill = nullptr; size_t mega_smem = 0; unsigned int *m_phase = nullptr; bf16 *m_fkv = nullptr; unsigned char *u_arena = nullptr; size_t u_arena_bytes = 0;
  dualar_tc *tc = nullptr; dualar_batch *batch = nullptr;
  int prefill_mode = 0;          // option prefill_mode: 0 = whole prompt through the tensor-core GEMMs, 1 = one position per launch (round-1 path, cross-check)
  int prefill_launches = 0;      // kernels the last tensor-core prefill launched
  bool prefix_reuse = true, kv_dirty = true; std::vector<int32_t> prev_prompt; int prev_T = 0, last_reuse = 0;   // KV reuse across requests (dualar_prefill)
  cudaEvent_t ev_ring[8] = {nullptr}; int ev_next = 0; int stream_cols = 0, steps_enqueued = 0;      // dualar_decode_async
  bool batch_keep_raw = true;    // batched decode keeps a copy of the raw logits for dualar_batch_read
  int batch_persistent = -1;     // batched decode as two persistent launches (bstep.cuh): -1 = DUALAR_BATCH_PERSIST or the default, 0 / 1 = option batch_persistent
  std::vector<dualar_batch *> groups; int group_slots = 32, batch_total = 0, batch_group_slots = 32; cudaEvent_t ev_groups_go = nullptr; cudaStream_t pf_stream = nullptr;      // request groups (batch_host.cuh)
  bool batch_decode_join = true; // dualar_batch_decode makes the caller's stream wait for every group's steps (option batch_decode_join)
  bool batch_fork = true;        // batched decode: LM head + slow sampler on a side stream beside fast pass 0 (DUALAR_BATCH_FORK=0: one stream)
  bool l2_window = false; float l2_hit_ratio = 0.0f; size_t l2_persist_bytes = 0;   // DUALAR_L2_WINDOW / DUALAR_L2_HIT: access-policy window over the fast stack
  bool chunk_group = true;   // DUALAR_CHUNK_GROUP=0: one 128-element chunk per unit everywhere (round 1 behaviour)
  bool use_mega = true;   // option mega_kernel / DUALAR_MEGA: 0 = one kernel per phase (the cross-check path)
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- arena planning ------------------------------------------------------------------------------
static void plan_layer(dualar_engine *e, const std::string &pre, LayerW &L, int dim, int nh, int nkv, int hd, int inter,
                       bool qkv_bias, bool o_bias, bool qk_norm, size_t &off, bool assign) {
  auto take = [&](bf16 *&ptr, const std::string &key, int rows, int cols, int pitch_rows, bf16 *alias_base, size_t alias_off) {
    if (alias_base) { if (assign) e->slots[key] = Slot{alias_base + alias_off, (int64_t)rows * cols, rows, cols, pitch_rows, false}; return; }
    size_t bytes = (size_t)rows * cols * pitch_rows * sizeof(bf16);
    if (assign) { ptr = (bf16 *)(e->arena + off); e->slots[key] = Slot{ptr, (int64_t)rows * cols, rows, cols, pitch_rows, false}; }
    off = align_up(off + bytes, 256);
  };
  take(L.wqkv, pre + ".attention.wqkv.weight", (nh + 2 * nkv) * hd, dim, 1, nullptr, 0);
  if (qkv_bias) take(L.bqkv, pre + ".attention.wqkv.bias", 1, (nh + 2 * nkv) * hd, 1, nullptr, 0);
  take(L.wo, pre + ".attention.wo.weight", dim, nh * hd, 1, nullptr, 0);
  if (o_bias) take(L.bo, pre + ".attention.wo.bias", 1, dim, 1, nullptr, 0);
  if (qk_norm) { take(L.qn, pre + ".attention.q_norm.weight", 1, hd, 1, nullptr, 0); take(L.kn, pre + ".attention.k_norm.weight", 1, hd, 1, nullptr, 0); }
  // w1 / w3 share one row-interleaved matrix: row 2j = w1[j], row 2j+1 = w3[j]
  take(L.w13, pre + ".feed_forward.w1.weight", inter, dim, 2, nullptr, 0);
  { bf16 *dummy = nullptr; take(dummy, pre + ".feed_forward.w3.weight", inter, dim, 2, assign ? L.w13 : (bf16 *)1, (size_t)dim); }
  take(L.w2, pre + ".feed_forward.w2.weight", dim, inter, 1, nullptr, 0);
  take(L.ffn_norm, pre + ".ffn_norm.weight", 1, dim, 1, nullptr, 0);
  take(L.attn_norm, pre + ".attention_norm.weight", 1, dim, 1, nullptr, 0);
}

static size_t plan(dualar_engine *e, bool assign) {
  const dualar_config &c = e->c;
  size_t off = 0;
  auto take = [&](bf16 *&ptr, const char *key, int64_t rows, int64_t cols) {
    size_t bytes = (size_t)rows * cols * sizeof(bf16);
    if (assign) { ptr = (bf16 *)(e->arena + off); e->slots[key] = Slot{ptr, rows * cols, (int)rows, (int)cols, 1, false}; }
    off = align_up(off + bytes, 256);
  };
  take(e->emb, "embeddings.weight", c.vocab_size, c.dim);
  take(e->cb_emb, "codebook_embeddings.weight", (int64_t)c.codebook_size * c.num_codebooks, c.dim);
  if (assign) e->slow.resize(c.n_layer);
  for (int i = 0; i < c.n_layer; ++i) {
    LayerW tmp; LayerW &L = assign ? e->slow[i] : tmp;
    plan_layer(e, "layers." + std::to_string(i), L, c.dim, c.n_head, c.n_local_heads, c.head_dim, c.intermediate_size,
               c.attention_qkv_bias, c.attention_o_bias, c.attention_qk_norm, off, assign);
  }
  take(e->norm, "norm.weight", 1, c.dim);
  if (!c.tie_word_embeddings) take(e->out_w, "output.weight", c.vocab_size, c.dim);
  take(e->rope, "freqs_cis", c.max_seq_len, c.head_dim);
  off = align_up(off, 2u << 20);
  if (assign) e->fast_off = off;
  if (c.fast_dim != c.dim) {      // fast_project_in = nn.Linear(dim, fast_dim) (llama.py:510-513, 590)
    take(e->fpi_w, "fast_project_in.weight", c.fast_dim, c.dim);
    take(e->fpi_b, "fast_project_in.bias", 1, c.fast_dim);
  }
  take(e->fast_emb, "fast_embeddings.weight", c.codebook_size, c.fast_dim);
  if (assign) e->fast.resize(c.n_fast_layer);
  for (int i = 0; i < c.n_fast_layer; ++i) {
    LayerW tmp; LayerW &L = assign ? e->fast[i] : tmp;
    plan_layer(e, "fast_layers." + std::to_string(i), L, c.fast_dim, c.fast_n_head, c.fast_n_local_heads, c.fast_head_dim,
               c.fast_intermediate_size, c.fast_attention_qkv_bias, c.fast_attention_o_bias, c.fast_attention_qk_norm, off, assign);
  }
  take(e->fast_norm, "fast_norm.weight", 1, c.fast_dim);
  take(e->fast_out, "fast_output.weight", c.codebook_size, c.fast_dim);
  take(e->fast_rope, "fast_freqs_cis", c.num_codebooks, c.fast_head_dim);
  if (assign) e->fast_bytes = off - e->fast_off;
  return align_up(off, 2u << 20);
}

static int check_config(const dualar_config &c) {
  if (c.abi_version != DUALAR_ABI_VERSION) return fail(DUALAR_EINVAL, "abi_version %d != %d", c.abi_version, DUALAR_ABI_VERSION);
  if (c.fast_dim % 256) return fail(DUALAR_EINVAL, "fast_dim must be a multiple of 256");
  if (c.dim % 256 || c.intermediate_size % 256 || (c.n_head * c.head_dim) % 256 || c.fast_intermediate_size % 256 ||
      (c.fast_n_head * c.fast_head_dim) % 256)
    return fail(DUALAR_EINVAL, "dim, intermediate_size and n_head*head_dim must be multiples of 256");
  if (c.head_dim % 8 || c.fast_head_dim % 8 || c.head_dim > 128 || c.head_dim < 8 || (c.head_dim & (c.head_dim - 1)))
    return fail(DUALAR_EINVAL, "head_dim must be a power of two in [8, 128]");
  if (c.n_head % c.n_local_heads || c.fast_n_head % c.fast_n_local_heads) return fail(DUALAR_EINVAL, "n_head %% n_local_heads != 0");
  int G = c.n_head / c.n_local_heads;
  if (G > DA_MAX_G || G * c.head_dim > DA_MAX_G * 128) return fail(DUALAR_EINVAL, "GQA group too large");
  if (c.n_local_heads > DA_MAX_KV_HEADS) return fail(DUALAR_EINVAL, "too many kv heads");
  if (c.num_codebooks + 1 > DA_MAX_ROWS || c.num_codebooks < 2) return fail(DUALAR_EINVAL, "num_codebooks out of range");
  if (c.fast_n_head * c.num_codebooks > 256 || c.num_codebooks * c.fast_n_local_heads * c.fast_head_dim > 8192 ||
      (c.fast_n_head + 2 * c.fast_n_local_heads) * c.fast_head_dim > 4096 || c.dim > 4096 || c.intermediate_size > 4096 ||
      c.fast_intermediate_size > 4096 || c.n_head * c.head_dim > 4096)
    return fail(DUALAR_EINVAL, "shape exceeds the per-CTA staging limits of the GEMV prologues (K <= 4096, fast attention <= 256 (head, position) pairs)");
  if (c.vocab_size <= 0 || c.max_seq_len <= 0 || c.n_layer <= 0 || c.n_fast_layer <= 0) return fail(DUALAR_EINVAL, "bad sizes");
  if (c.semantic_begin_id < 0 || c.semantic_end_id >= c.vocab_size || c.semantic_end_id < c.semantic_begin_id)
    return fail(DUALAR_EINVAL, "semantic id range outside the vocabulary");
  return 0;
}

extern "C" int dualar_abi_version(void) { return DUALAR_ABI_VERSION; }
extern "C" const char *dualar_last_error(void) { return g_err; }

extern "C" int dualar_create(const dualar_config *cfg, int device, dualar_engine **out) {
  if (!cfg || !out) return fail(DUALAR_EINVAL, "null argument");
  int rc = check_config(*cfg); if (rc) return rc;
  int ndev = 0; CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(DUALAR_EINVAL, "device %d not present (%d visible)", device, ndev);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(DUALAR_EINVAL, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  dualar_engine *e = new dualar_engine();
  e->c = *cfg; e->device = device; e->sms = prop.multiProcessorCount;
  e->fv = cfg->codebook_size < 1024 ? cfg->codebook_size : 1024;   // logits[:, :, :1024]  (inference.py:134)
  e->arena_bytes = plan(e, false);
  cudaError_t ce = cudaMalloc((void **)&e->arena, e->arena_bytes);
  if (ce != cudaSuccess) { delete e; return fail(DUALAR_ECUDA, "cudaMalloc(%zu) failed: %s", e->arena_bytes, cudaGetErrorString(ce)); }
  cudaMemset(e->arena, 0, e->arena_bytes);
  plan(e, true);
  *out = e;
  return 0;
}

extern "C" int dualar_load_weight(dualar_engine *e, const char *key, const void *data, int64_t n, int on_device) {
  if (!e || !key || !data) return fail(DUALAR_EINVAL, "null argument");
  if (e->finalized) return fail(DUALAR_ESTATE, "weights are frozen after dualar_finalize");
  auto it = e->slots.find(key);
  if (it == e->slots.end()) return fail(DUALAR_EINVAL, "unknown weight key '%s'", key);
  Slot &s = it->second;
  if (n != s.n) return fail(DUALAR_EINVAL, "weight '%s': %lld elements given, %lld expected", key, (long long)n, (long long)s.n);
  CU(cudaSetDevice(e->device));
  cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  size_t row = (size_t)s.cols * sizeof(bf16);
  if (s.pitch_rows == 1) CU(cudaMemcpy(s.dst, data, (size_t)s.n * sizeof(bf16), kind));
  else CU(cudaMemcpy2D(s.dst, row * s.pitch_rows, data, row, row, s.rows, kind));
  s.loaded = true;
  if (!strcmp(key, "freqs_cis")) e->rope_loaded = true;
  if (!strcmp(key, "fast_freqs_cis")) e->fast_rope_loaded = true;
  return 0;
}

extern "C" int dualar_bind_kv(dualar_engine *e, int is_fast, int layer, void *k, void *v) {
  if (!e || !k || !v) return fail(DUALAR_EINVAL, "null argument");
  if (e->finalized) return fail(DUALAR_ESTATE, "bind KV before dualar_finalize");
  std::vector<LayerW> &Ls = is_fast ? e->fast : e->slow;
  if (layer < 0 || layer >= (int)Ls.size()) return fail(DUALAR_EINVAL, "layer %d out of range", layer);
  Ls[layer].kc = (bf16 *)k; Ls[layer].vc = (bf16 *)v; Ls[layer].kv_external = true;
  return 0;
}

// llama.py:594-603 restated on the host (used only when the caller does not hand the table over)
static void host_rope(std::vector<uint16_t> &out, int seq, int n_elem, float base) {
  out.resize((size_t)seq * n_elem);
  for (int i = 0; i < n_elem / 2; ++i) {
    float ex = (float)(2 * i) / (float)n_elem;
    float freq = 1.0f / (float)pow((double)base, (double)ex);
    for (int t = 0; t < seq; ++t) {
      float ang = (float)t * freq;
      float cs[2] = {(float)cos((double)ang), (float)sin((double)ang)};
      for (int k = 0; k < 2; ++k) {
        uint32_t u; memcpy(&u, &cs[k], 4);
        uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
        out[((size_t)t * (n_elem / 2) + i) * 2 + k] = (uint16_t)(r >> 16);
      }
    }
  }
}

// every kernel of the step graph is launched with programmatic stream serialization (PDL): it starts while its
// predecessor still runs, prefetches weights, and blocks in griddepcontrol.wait until the predecessor has finished
static bool g_use_pdl = true;
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

template <typename T> static int dev_alloc(dualar_engine *e, T *&p, size_t n) {
  CU(cudaMalloc((void **)&p, n * sizeof(T)));
  CU(cudaMemset(p, 0, n * sizeof(T)));
  e->owned.push_back((void *)p);
  return 0;
}

// ---- kernel launch plumbing ---------------------------------------------------------------------------
template <int PRO, int EPI> static size_t gemv_smem(const dualar_engine *e, const GemvArgs &a) {
  size_t f = ((size_t)a.K + 80) * sizeof(float);
  size_t work = 0;
  if (PRO == PRO_FASTATTN) work = ((size_t)a.fa.nh * a.fa.hd + 2 * (size_t)a.fa.ncb * a.fa.nkv * a.fa.hd + (size_t)a.fa.nh * a.fa.ncb) * sizeof(float);
  if (EPI == EPI_LOGITS && a.head > 0) { size_t w2 = 192 * 8 + 80 * 4 + 64; if (w2 > work) work = w2; }
  (void)e;
  return f + work + 16;
}
template <int PRO, int EPI> static int launch_gemv(dualar_engine *e, GemvArgs a, cudaStream_t s, int &count) {
  static size_t configured[64] = {0};      // function attributes are per device: one high-water mark per device ordinal
  size_t smem = gemv_smem<PRO, EPI>(e, a);
  size_t &conf = configured[e->device & 63];
  if (smem > conf) {
    CU(cudaFuncSetAttribute(gemv_kernel<PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 200 * 1024 ? smem : 200 * 1024)));
    conf = 200 * 1024 > smem ? 200 * 1024 : smem;
  }
  int npairs = (a.rows + 1) / 2;
  int grid = npairs < 2 * e->sms ? npairs : 2 * e->sms;
  a.st = e->st; a.tl.buf = e->tl; a.tl.slot = count;
  CU(launch_k(gemv_kernel<PRO, EPI>, dim3(grid), dim3(DA_GEMV_THREADS), smem, s, a));
  ++count;
  return grid;
}

static GemvArgs base_args(const bf16 *W, const bf16 *bias, int rows, int K, int evict_last) {
  GemvArgs a; memset(&a, 0, sizeof(a));
  a.W = W; a.bias = bias; a.rows = rows; a.K = K; a.evict_last = evict_last;
  return a;
}

// one transformer block on the slow stack (llama.py:322-331)
static int enqueue_slow_layer(dualar_engine *e, int li, cudaStream_t s, int &count) {
  const dualar_config &c = e->c; LayerW &L = e->slow[li];
  const int qkv_rows = (c.n_head + 2 * c.n_local_heads) * c.head_dim, qd = c.n_head * c.head_dim;
  int rc;
  { GemvArgs a = base_args(L.wqkv, L.bqkv, qkv_rows, c.dim, 0); a.x = e->x; a.norm_w = L.attn_norm; a.eps = c.norm_eps; a.out = e->qkv;
    if ((rc = launch_gemv<PRO_RMSNORM, EPI_STORE>(e, a, s, count)) < 0) return rc; }
  { AttnArgs t; memset(&t, 0, sizeof(t));
    t.qkv = e->qkv; t.kc = L.kc; t.vc = L.vc; t.rope = e->rope; t.qn = L.qn; t.kn = L.kn;
    t.nh = c.n_head; t.nkv = c.n_local_heads; t.hd = c.head_dim; t.S = c.max_seq_len; t.eps = c.norm_eps;
    t.sf = (float)sqrt(1.0 / sqrt((double)c.head_dim));
    t.part_o = e->part_o; t.part_ml = e->part_ml; t.y = e->y; t.nsplit_max = e->nsplit; t.st = e->st; t.tl.buf = e->tl; t.tl.slot = count;
    size_t smem = attn_smem_bytes(c.n_head / c.n_local_heads, c.head_dim);
    static size_t configured[64] = {0};   // per device: grow the opt-in limit when a bigger shape comes along
    if (smem > configured[e->device & 63]) { CU(cudaFuncSetAttribute(attn_slow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured[e->device & 63] = smem; }
    CU(launch_k(attn_slow_kernel, dim3(e->nsplit, c.n_local_heads), dim3(DA_ATTN_THREADS), smem, s, t)); ++count; }
  { GemvArgs a = base_args(L.wo, L.bo, c.dim, qd, 0); a.x = e->y; a.res = e->x; a.out = e->h;
    if ((rc = launch_gemv<PRO_PLAIN, EPI_RESIDUAL>(e, a, s, count)) < 0) return rc; }
  { GemvArgs a = base_args(L.w13, nullptr, 2 * c.intermediate_size, c.dim, 0); a.x = e->h; a.norm_w = L.ffn_norm; a.eps = c.norm_eps; a.out = e->act;
    if ((rc = launch_gemv<PRO_RMSNORM, EPI_SWIGLU>(e, a, s, count)) < 0) return rc; }
  { GemvArgs a = base_args(L.w2, nullptr, c.dim, c.intermediate_size, 0); a.x = e->act; a.res = e->h; a.out = e->x;
    if ((rc = launch_gemv<PRO_PLAIN, EPI_RESIDUAL>(e, a, s, count)) < 0) return rc; }
  return 0;
}

// one fast block for codebook position `p` (llama.py:561-580 via :322-331, use_sdpa=False)
static int enqueue_fast_layer(dualar_engine *e, int li, int p, const bf16 *in, bf16 *out, cudaStream_t s, int &count) {
  const dualar_config &c = e->c; LayerW &L = e->fast[li];
  const int qkv_rows = (c.fast_n_head + 2 * c.fast_n_local_heads) * c.fast_head_dim, qd = c.fast_n_head * c.fast_head_dim;
  int rc;
  { GemvArgs a = base_args(L.wqkv, L.bqkv, qkv_rows, c.fast_dim, 1); a.x = in; a.norm_w = L.attn_norm; a.eps = c.norm_eps; a.out = e->fqkv;
    if ((rc = launch_gemv<PRO_RMSNORM, EPI_STORE>(e, a, s, count)) < 0) return rc; }
  { GemvArgs a = base_args(L.wo, L.bo, c.fast_dim, qd, 1); a.res = in; a.out = e->fh;
    a.fa.qkv = e->fqkv; a.fa.kc = L.kc; a.fa.vc = L.vc; a.fa.rope = e->fast_rope; a.fa.qn = L.qn; a.fa.kn = L.kn;
    a.fa.nh = c.fast_n_head; a.fa.nkv = c.fast_n_local_heads; a.fa.hd = c.fast_head_dim; a.fa.ncb = c.num_codebooks; a.fa.pos = p;
    a.fa.eps = c.norm_eps; a.fa.scale = (float)(1.0 / sqrt((double)c.fast_head_dim));
    if ((rc = launch_gemv<PRO_FASTATTN, EPI_RESIDUAL>(e, a, s, count)) < 0) return rc; }
  { GemvArgs a = base_args(L.w13, nullptr, 2 * c.fast_intermediate_size, c.fast_dim, 1); a.x = e->fh; a.norm_w = L.ffn_norm; a.eps = c.norm_eps; a.out = e->fact;
    if ((rc = launch_gemv<PRO_RMSNORM, EPI_SWIGLU>(e, a, s, count)) < 0) return rc; }
  { GemvArgs a = base_args(L.w2, nullptr, c.fast_dim, c.fast_intermediate_size, 1); a.x = e->fact; a.res = e->fh; a.out = out;
    if ((rc = launch_gemv<PRO_PLAIN, EPI_RESIDUAL>(e, a, s, count)) < 0) return rc; }
  return 0;
}

// the decode step of decode_one_token_ar (inference.py:83-155); slow_only = one prefill position
static int enqueue_step(dualar_engine *e, cudaStream_t s, bool slow_only, int &count) {
  const dualar_config &c = e->c;
  int rc;
  if (e->use_mega) {
    // the whole step as ONE persistent cooperative kernel (mega.cuh); the phase tables were built at finalize
    MegaArgs *a = slow_only ? e->ma_prefill : e->ma_step;
    a->tl = e->tl; a->tl_slots = e->tl_slots; a->tl2 = e->tl2;
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(e->sms); cfg.blockDim = dim3(DA_M_THREADS); cfg.dynamicSmemBytes = e->mega_smem; cfg.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const bool tl_on = e->tl != nullptr;
    if (slow_only) { if (tl_on) CU(cudaLaunchKernelEx(&cfg, mega_kernel<true, false>, *a)); else CU(cudaLaunchKernelEx(&cfg, mega_kernel<false, false>, *a)); }
    else { if (tl_on) CU(cudaLaunchKernelEx(&cfg, mega_kernel<true, true>, *a)); else CU(cudaLaunchKernelEx(&cfg, mega_kernel<false, true>, *a)); }
    ++count;
    return 0;
  }
  { EmbedArgs a; memset(&a, 0, sizeof(a));
    a.emb = e->emb; a.cb_emb = e->cb_emb; a.x = e->x; a.dim = c.dim; a.vocab = c.vocab_size; a.codebook_size = c.codebook_size;
    a.num_codebooks = c.num_codebooks; a.sem_begin = c.semantic_begin_id; a.sem_end = c.semantic_end_id; a.scale_cb = c.scale_codebook_embeddings;
    a.inv_sqrt = (float)(1.0 / sqrt((double)(c.num_codebooks + 1))); a.sqrt_c = (float)sqrt((double)(c.num_codebooks + 1)); a.st = e->st; a.tl.buf = e->tl; a.tl.slot = count;
    CU(launch_k(embed_kernel, dim3((c.dim + 255) / 256), dim3(256), 0, s, a)); ++count; }
  for (int i = 0; i < c.n_layer; ++i) if ((rc = enqueue_slow_layer(e, i, s, count)) < 0) return rc;
  if (slow_only) {
    PrefillColArgs a{e->seq, c.max_seq_len, c.num_codebooks + 1, e->st};
    CU(launch_k(prefill_col_kernel, dim3(1), dim3(32), 0, s, a, 1)); ++count;
    return 0;
  }
  // LM head + sampler (llama.py:446-451, inference.py:103-113)
  int head_grid;
  { GemvArgs a = base_args(c.tie_word_embeddings ? e->emb : e->out_w, nullptr, c.vocab_size, c.dim, 0);
    a.x = e->x; a.norm_w = e->norm; a.eps = c.norm_eps; a.out = e->logits; a.logits_raw = e->logits_raw; a.partials = e->partials;
    a.head = 0; a.n_rows_tok = c.num_codebooks + 1;
    if ((head_grid = launch_gemv<PRO_RMSNORM, EPI_LOGITS>(e, a, s, count)) < 0) return head_grid; }
  { SelectArgs a; memset(&a, 0, sizeof(a));
    a.logits = e->logits; a.partials = e->partials; a.n_partials = head_grid; a.V = c.vocab_size; a.delta = e->delta; a.cand = e->cand;
    a.fast_emb = e->fast_emb; a.fast_x = e->fin; a.fast_dim = c.fast_dim; a.codebook_size = c.codebook_size; a.sem_begin = c.semantic_begin_id; a.st = e->st; a.tl.buf = e->tl; a.tl.slot = count;
    size_t smem = 192 * 8 + 34 * 8 + 80 * 4 + 64;
    static bool configured[64] = {false};   // per device
    if (!configured[e->device & 63]) { CU(cudaFuncSetAttribute(select_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured[e->device & 63] = true; }
    CU(launch_k(select_sample_kernel, dim3(e->sms), dim3(512), smem, s, a)); ++count; }
  // fast AR: pass 0 consumes the slow hidden state, pass k >= 1 the embedding of codebook k-1 (inference.py:121-149)
  if (e->fpi_w) {      // hidden_states = fast_project_in(x)   (llama.py:590)
    GemvArgs a = base_args(e->fpi_w, e->fpi_b, c.fast_dim, c.dim, 1); a.x = e->x; a.out = e->fpi;
    if ((rc = launch_gemv<PRO_PLAIN, EPI_STORE>(e, a, s, count)) < 0) return rc;
  }
  for (int p = 0; p < c.num_codebooks; ++p) {
    const bf16 *in = p == 0 ? (e->fpi_w ? e->fpi : e->x) : e->fin;
    for (int l = 0; l < c.n_fast_layer; ++l) {
      bf16 *out = e->fbuf[l & 1];
      if ((rc = enqueue_fast_layer(e, l, p, in, out, s, count)) < 0) return rc;
      in = out;
    }
    if (p == 0) continue;   // logits of pass 0 are discarded by the reference (inference.py:122)
    GemvArgs a = base_args(e->fast_out, nullptr, e->fv, c.fast_dim, 1);
    a.x = in; a.norm_w = e->fast_norm; a.eps = c.norm_eps;
    a.out = e->flogits + (size_t)(p - 1) * e->fv; a.logits_raw = e->flogits_raw + (size_t)(p - 1) * e->fv;
    a.head = p; a.noise_off = (long long)c.vocab_size + (long long)(p - 1) * e->fv;
    a.fast_emb = e->fast_emb; a.fast_x = e->fin; a.fast_dim = c.fast_dim; a.codebook_size = c.codebook_size;
    a.last_head = (p == c.num_codebooks - 1);
    a.seq = e->seq; a.seq_stride = c.max_seq_len; a.im_end_id = c.im_end_id; a.n_rows_tok = c.num_codebooks + 1;
    if ((rc = launch_gemv<PRO_RMSNORM, EPI_LOGITS>(e, a, s, count)) < 0) return rc;
  }
  return 0;
}

static int capture(dualar_engine *e, bool slow_only, cudaGraphExec_t *out, int *count) {
  // dry run first: configures kernel attributes outside capture and surfaces launch errors eagerly
  int n = 0;
  int rc = enqueue_step(e, e->cap_stream, slow_only, n);
  if (rc < 0) return rc;
  CU(cudaStreamSynchronize(e->cap_stream));
  cudaGraph_t g;
  CU(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
  n = 0;
  rc = enqueue_step(e, e->cap_stream, slow_only, n);
  cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &g);
  if (rc < 0) return rc;
  if (ce != cudaSuccess) return fail(DUALAR_ECUDA, "graph capture failed: %s", cudaGetErrorString(ce));
  if (e->l2_window && !slow_only) {
    // the fast stack is re-read num_codebooks times per token: cover it with an access-policy window on every kernel node of the
    // decode graph (hitRatio = the share of the window the persisting carve-out can hold), everything else streams
    size_t nn = 0; CU(cudaGraphGetNodes(g, nullptr, &nn));
    std::vector<cudaGraphNode_t> nodes(nn); CU(cudaGraphGetNodes(g, nodes.data(), &nn));
    int max_win = 0; cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, e->device);
    // the hot set: the fast layers and the rows of fast_output that are used (fast_embeddings ahead of them is a row gather)
    char *w0 = (char *)e->fast[0].wqkv, *w1 = (char *)(e->fast_out + (size_t)e->fv * e->c.fast_dim);
    size_t bytes = (size_t)(w1 - w0); if (max_win > 0 && bytes > (size_t)max_win) bytes = (size_t)max_win;
    for (auto nd : nodes) {
      cudaGraphNodeType ty; CU(cudaGraphNodeGetType(nd, &ty));
      if (ty != cudaGraphNodeTypeKernel) continue;
      cudaKernelNodeAttrValue v; memset(&v, 0, sizeof(v));
      v.accessPolicyWindow.base_ptr = w0; v.accessPolicyWindow.num_bytes = bytes;
      v.accessPolicyWindow.hitRatio = e->l2_hit_ratio > 0.f ? e->l2_hit_ratio : (float)((double)e->l2_persist_bytes / (double)bytes > 1.0 ? 1.0 : (double)e->l2_persist_bytes / (double)bytes); v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      CU(cudaGraphKernelNodeSetAttribute(nd, cudaKernelNodeAttributeAccessPolicyWindow, &v));
    }
  }
  CU(cudaGraphInstantiate(out, g, 0));
  CU(cudaGraphDestroy(g));
  *count = n;
  return 0;
}

// ---- the static schedule of the persistent whole-step kernel (mega.cuh) ------------------------------------------------
// Falls back to one kernel per phase (use_mega = false) when a shape does not fit the kernel's limits.
static int build_mega(dualar_engine *e) {
  const dualar_config &c = e->c;
  const int grid = e->sms;
  const int qd = c.n_head * c.head_dim, kd = c.n_local_heads * c.head_dim, qkv_rows = qd + 2 * kd;
  const int fqd = c.fast_n_head * c.fast_head_dim, fkd = c.fast_n_local_heads * c.fast_head_dim, fqkv_rows = fqd + 2 * fkd;
  const int G = c.n_head / c.n_local_heads;
  const int n_slow = 6 * c.n_layer, n_step = n_slow + 3 + 4 * c.n_fast_layer + c.num_codebooks * 4 * c.n_fast_layer + (c.num_codebooks - 1);
  auto pairs = [](int rows) { return (rows + 1) / 2; };
  if (n_step > DA_M_MAX_PHASES || c.n_layer > DA_M_MAXL || c.n_fast_layer > DA_M_MAXFL || grid > 160 || e->fv > 1024 ||
      pairs(qkv_rows) < grid || pairs(2 * c.intermediate_size) < grid || pairs(fqkv_rows) < grid || pairs(2 * c.fast_intermediate_size) < grid ||
      c.fast_dim > 8 * DA_M_CTHREADS ||
      G * c.head_dim > 1024 || c.vocab_size > (1 << 18) || (c.num_codebooks - 1) * 2 * fkd / 8 > 3 * DA_M_CTHREADS || c.head_dim < 32 ||
      c.dim > 8 * DA_M_CTHREADS || c.intermediate_size > 8 * DA_M_CTHREADS || c.fast_intermediate_size > 8 * DA_M_CTHREADS || qd > 8 * DA_M_CTHREADS || fqkv_rows > 8 * DA_M_CTHREADS ||
      3 * grid > DA_M_CTHREADS ||
      (c.head_dim & (c.head_dim - 1)) || c.head_dim > 256 ||                                            // a head vector = hd/8 lanes of one warp (attention staging)
      (c.fast_head_dim & (c.fast_head_dim - 1)) || c.fast_head_dim < 32 || c.fast_head_dim > 128 ||     // ... hd/4 lanes in the fast stack
      ((c.fast_n_head + 2 * c.fast_n_local_heads) & 1)) { e->use_mega = false; return 0; }
  int rc;
  // every broadcast unit vector exists DA_M_REP times, `ustride` units apart
  int umax = qkv_rows; for (int v : {c.dim, qd, c.intermediate_size, fqkv_rows, c.fast_dim, c.fast_intermediate_size, e->fv}) if (v > umax) umax = v;
  const size_t ustride = align_up((size_t)umax, 256) + 64, ubuf = ustride * DA_M_REP;
  // one arena for every tagged buffer, so that a request can start from "no tag is valid" with a single memset:
  // tags are 16 bits wide, and a buffer that only the decode kernel writes would otherwise keep the previous request's
  // units across the prefill launches in between (a tag collision there would hand a consumer stale data)
  {
    const size_t n_po = (size_t)c.n_local_heads * e->nsplit * G * c.head_dim, n_pml = (size_t)c.n_local_heads * e->nsplit * G * 2;
    size_t bytes = 13 * ubuf * 4 + (n_po + n_pml + (size_t)grid + 3 * (size_t)grid + DA_CAND_CAP) * 8;
    if ((rc = dev_alloc(e, e->u_arena, bytes))) return rc;
    e->u_arena_bytes = bytes;
    unsigned long long *q = (unsigned long long *)e->u_arena;      // 64-bit buffers first (alignment)
    e->m_part_o = q; q += n_po; e->m_part_ml = q; q += n_pml; e->m_hmax = q; q += grid; e->m_hcs = q; q += 3 * grid; e->m_cand = q; q += DA_CAND_CAP;
    uint32_t *u = (uint32_t *)q;
    uint32_t **slots[13] = {&e->u_x, &e->u_qkv, &e->u_y, &e->u_h, &e->u_act, &e->u_fqkv, &e->u_fh, &e->u_fact, &e->u_fx0, &e->u_fx1, &e->u_fin, &e->u_flogits, &e->u_fpi};
    for (auto sl : slots) { *sl = u; u += ubuf; }
  }
  if ((rc = dev_alloc(e, e->m_phase, 1))) return rc;
  if ((rc = dev_alloc(e, e->m_fkv, (size_t)grid * c.n_fast_layer * c.num_codebooks * 2 * fkd))) return rc;
  { const char *v = getenv("DUALAR_T0"); if (v) e->use_t0 = v[0] != '0'; }
  { const char *v = getenv("DUALAR_CHUNK_GROUP"); if (v) e->chunk_group = v[0] != '0'; }
  if (e->use_t0) {
    // table of the first fast layer's q | k | v for every code (see t0_build_kernel); 16 MB for s1-mini
    if ((rc = dev_alloc(e, e->t0, (size_t)c.codebook_size * fqkv_rows))) return rc;
    t0_build_kernel<<<c.codebook_size, 256, (size_t)c.fast_dim * sizeof(float)>>>(e->fast_emb, e->fast[0].attn_norm, e->fast[0].wqkv, e->fast[0].bqkv, e->t0,
                                                                                   c.fast_dim, fqkv_rows, c.norm_eps);
    CU(cudaGetLastError());
  }
  e->ma_step = new MegaArgs(); e->ma_prefill = new MegaArgs();
  for (int variant = 0; variant < 2; ++variant) {
    MegaArgs &a = variant ? *e->ma_prefill : *e->ma_step;
    memset(&a, 0, sizeof(a));
    std::vector<MPhase> tab;
    auto gemv = [&](const bf16 *W, const bf16 *bias, const bf16 *norm_w, const uint32_t *in, int in_ph, uint32_t *out, int rows, int K,
                    int pro, int epi, int flags, int layer, int pos) {
      MPhase p; memset(&p, 0, sizeof(p));
      p.W = W; p.bias = bias; p.norm_w = norm_w; p.in = in; p.out = out; p.rows = rows; p.K = K; p.in_ph = (short)in_ph;
      p.pq = (short)(((rows + 1) / 2) / grid); p.prem = (short)(((rows + 1) / 2) % grid);
      p.kind = MK_GEMV; p.pro = (unsigned char)pro; p.epi = (unsigned char)epi; p.flags = (unsigned char)flags; p.layer = (unsigned char)layer; p.pos = (unsigned char)pos;
      // chunks per unit: with more than 16 (tile, chunk) units on the busiest CTA half the warps would take a second round; a warp then
      // computes `cgroup` consecutive chunks per unit instead (the per-chunk partials and their summation order stay what they were)
      { const int nchunk = K >> 7, tiles = (2 * (((rows + 1) / 2) / grid + 1) + 15) / 16;
        int cg = 1;
        if (e->chunk_group) while (tiles * nchunk / cg > DA_M_CWARPS && cg < nchunk && cg < 4) { ++cg; while (nchunk % cg) ++cg; }
        p.cgroup = (unsigned char)cg; }
      tab.push_back(p); return (int)tab.size() - 1;
    };
    auto other = [&](int kind, const uint32_t *in, int in_ph, uint32_t *out, int layer) {
      MPhase p; memset(&p, 0, sizeof(p));
      p.kind = (unsigned char)kind; p.in = in; p.in_ph = (short)in_ph; p.out = out; p.layer = (unsigned char)layer;
      tab.push_back(p); return (int)tab.size() - 1;
    };
    int last = 0;
    for (int l = 0; l < c.n_layer; ++l) {
      LayerW &L = e->slow[l];
      int q = gemv(L.wqkv, L.bqkv, L.attn_norm, l ? e->u_x : nullptr, last, e->u_qkv, qkv_rows, c.dim, l ? MP_RMSNORM : MP_EMBED, ME_STORE, MF_SAVE0, l, 0);
      int at = other(MK_ATTN, e->u_qkv, q, nullptr, l);
      int mg = other(MK_MERGE, nullptr, at, e->u_y, l);
      int o = gemv(L.wo, L.bo, nullptr, e->u_y, mg, e->u_h, c.dim, qd, MP_PLAIN, ME_RESIDUAL, MF_RES0, l, 0);
      int f = gemv(L.w13, nullptr, L.ffn_norm, e->u_h, o, e->u_act, 2 * c.intermediate_size, c.dim, MP_RMSNORM, ME_SWIGLU, MF_SAVE1, l, 0);
      last = gemv(L.w2, nullptr, nullptr, e->u_act, f, e->u_x, c.dim, c.intermediate_size, MP_PLAIN, ME_RESIDUAL, MF_RES1, l, 0);
    }
    if (variant) other(MK_PREFILL_END, e->u_x, last, nullptr, 0);
    else {
      // The LM head (319 MB, HBM bound, no dependency but x) and the fast pass 0 (16 latency-bound phases that need only x) are
      // independent: the head comes in FL*4 parts, one after each pass-0 phase, so the hand-over latency of a pass-0 phase
      // hides behind a slice of the head's tiles instead of idling the machine.
      const int FL = c.n_fast_layer, NPARTS = 4 * FL;
      int hd_ph = -1, prev = 0, part = 0;
      auto head_part = [&]() {
        hd_ph = gemv(c.tie_word_embeddings ? e->emb : e->out_w, nullptr, e->norm, e->u_x, last, nullptr, c.vocab_size, c.dim, MP_RMSNORM, ME_SLOWLOGITS, 0, 0, 0);
        tab.back().part = (unsigned char)part; tab.back().nparts = (unsigned char)NPARTS; ++part;
        { const int nchunk = c.dim >> 7, tiles_all = (2 * (((c.vocab_size + 1) / 2) / grid + 1) + 15) / 16, tiles = (tiles_all + NPARTS - 1) / NPARTS;
          int cg = 1;
          if (e->chunk_group) while (tiles * nchunk / cg > DA_M_CWARPS && cg < nchunk && cg < 4) { ++cg; while (nchunk % cg) ++cg; }
          tab.back().cgroup = (unsigned char)cg; }
      };
      auto fast_layer = [&](int p, int l, const uint32_t *lin, int lin_ph, bool with_head) {
        LayerW &W = e->fast[l];
        if (with_head) head_part();
        // passes >= 1, first layer: the input is the embedding of a code, so wqkv's product comes from the table and the phase
        // is dropped; the attention staging of wo reads the code (one unit in u_fin), the table row and the embedding row
        const bool tab = e->t0 && p >= 1 && l == 0;
        int q = tab ? lin_ph : gemv(W.wqkv, W.bqkv, W.attn_norm, lin, lin_ph, e->u_fqkv, fqkv_rows, c.fast_dim, MP_RMSNORM, ME_STORE, MF_KEEP | MF_SAVE0, l, p);
        if (with_head) head_part();
        int o = gemv(W.wo, W.bo, nullptr, tab ? lin : e->u_fqkv, q, e->u_fh, c.fast_dim, fqd, MP_FASTATTN, ME_RESIDUAL, MF_KEEP | MF_RES0 | (tab ? MF_T0 : 0), l, p);
        if (with_head) head_part();
        int f = gemv(W.w13, nullptr, W.ffn_norm, e->u_fh, o, e->u_fact, 2 * c.fast_intermediate_size, c.fast_dim, MP_RMSNORM, ME_SWIGLU, MF_KEEP | MF_SAVE1, l, p);
        if (with_head) head_part();
        return gemv(W.w2, nullptr, nullptr, e->u_fact, f, (l & 1) ? e->u_fx1 : e->u_fx0, c.fast_dim, c.fast_intermediate_size, MP_PLAIN, ME_RESIDUAL, MF_KEEP | MF_RES1, l, p);
      };
      // fast_dim != dim: pass 0 starts from fast_project_in(x) (llama.py:590), one more GEMV phase (bias in the epilogue)
      int p0_ph = last; const uint32_t *p0_in = e->u_x;
      if (e->fpi_w) { p0_ph = gemv(e->fpi_w, e->fpi_b, nullptr, e->u_x, last, e->u_fpi, c.fast_dim, c.dim, MP_PLAIN, ME_STORE, MF_KEEP, 0, 0); p0_in = e->u_fpi; }
      for (int l = 0; l < FL; ++l) {      // pass 0 (input: the slow hidden state), interleaved with the head
        const uint32_t *lin = l == 0 ? p0_in : (((l - 1) & 1) ? e->u_fx1 : e->u_fx0);
        prev = fast_layer(0, l, lin, l == 0 ? p0_ph : prev, true);
      }
      int hs = other(MK_HSTAT, nullptr, hd_ph, nullptr, 0);
      int hc = other(MK_HCAND, nullptr, hs, nullptr, 0);
      int fin_ph = hc;   // the phase that last wrote u_fin
      for (int p = 1; p < c.num_codebooks; ++p) {
        for (int l = 0; l < FL; ++l) {
          const uint32_t *lin = l == 0 ? e->u_fin : (((l - 1) & 1) ? e->u_fx1 : e->u_fx0);
          prev = fast_layer(p, l, lin, l == 0 ? fin_ph : prev, false);
        }
        fin_ph = gemv(e->fast_out, nullptr, e->fast_norm, ((FL - 1) & 1) ? e->u_fx1 : e->u_fx0, prev, e->u_flogits, e->fv, c.fast_dim, MP_RMSNORM, ME_FASTLOGITS, MF_KEEP, FL - 1, p);
      }
    }
    a.n_phases = (int)tab.size();
    memcpy(a.table, tab.data(), tab.size() * sizeof(MPhase));
    a.rope = e->rope; a.nh = c.n_head; a.nkv = c.n_local_heads; a.hd = c.head_dim; a.S = c.max_seq_len; a.nsplit_max = e->nsplit;
    a.eps = c.norm_eps; a.sf = (float)sqrt(1.0 / sqrt((double)c.head_dim));
    for (int l = 0; l < c.n_layer; ++l) { a.kc[l] = e->slow[l].kc; a.vc[l] = e->slow[l].vc; a.qn[l] = e->slow[l].qn; a.kn[l] = e->slow[l].kn; }
    a.part_o = e->m_part_o; a.part_ml = e->m_part_ml;
    a.emb = e->emb; a.cb_emb = e->cb_emb; a.dim = c.dim; a.vocab = c.vocab_size; a.codebook_size = c.codebook_size; a.num_codebooks = c.num_codebooks;
    a.sem_begin = c.semantic_begin_id; a.sem_end = c.semantic_end_id; a.scale_cb = c.scale_codebook_embeddings;
    a.inv_sqrt = (float)(1.0 / sqrt((double)(c.num_codebooks + 1))); a.sqrt_c = (float)sqrt((double)(c.num_codebooks + 1));
    a.logits = e->logits; a.logits_raw = e->logits_raw; a.hmax = e->m_hmax; a.hcs = e->m_hcs; a.cand = e->m_cand; a.delta = e->delta; a.n_rows_tok = c.num_codebooks + 1;
    a.head_pq = ((c.vocab_size + 1) / 2) / grid; a.head_prem = ((c.vocab_size + 1) / 2) % grid;
    a.frope = e->fast_rope;
    for (int l = 0; l < c.n_fast_layer; ++l) { a.fqn[l] = e->fast[l].qn; a.fkn[l] = e->fast[l].kn; }
    a.fl = c.n_fast_layer; a.fnh = c.fast_n_head; a.fnkv = c.fast_n_local_heads; a.fhd = c.fast_head_dim; a.ncb = c.num_codebooks;
    a.fscale = (float)(1.0 / sqrt((double)c.fast_head_dim));
    a.fkv = e->m_fkv; a.t0 = e->t0; a.fast_emb = e->fast_emb; a.fdim = c.fast_dim; a.fv = e->fv; a.u_fin = e->u_fin; a.flogits_raw = e->flogits_raw; a.flogits = e->flogits;
    a.noise_off0 = (long long)c.vocab_size;
    a.seq = e->seq; a.seq_stride = c.max_seq_len; a.im_end_id = c.im_end_id; a.st = e->st; a.phase_ctr = e->m_phase;
    // shared-memory plan (from the decode-step table; the prefill table is a subset and shares it)
    if (variant) {
      const MegaArgs &s0 = *e->ma_step;
      a.kmax = s0.kmax; a.lg_rows = s0.lg_rows; a.work_bytes = s0.work_bytes; a.kv_bytes = s0.kv_bytes; a.ring_bytes = s0.ring_bytes; a.plan = s0.plan;
      continue;
    }
    int kmax = 0; size_t max_entry = 0;
    for (auto &p : tab) if (p.kind == MK_GEMV) {
      if (p.K > kmax) kmax = p.K;
      int nr_max = 2 * (pairs(p.rows) / grid + 1); if (nr_max > 16) nr_max = 16;      // rows of the largest tile of any CTA
      const size_t entry = (size_t)nr_max * (2 * (size_t)p.K + 16);
      if (entry > max_entry) max_entry = entry;
    }
    if ((size_t)4 * DA_TILE * c.head_dim > max_entry) max_entry = (size_t)4 * DA_TILE * c.head_dim;
    size_t work = (size_t)(fqd + 2 * fkd + c.fast_n_head * 16 + 4) * 4;
    auto upd = [&](size_t v) { if (v > work) work = v; };
    upd((size_t)(G * c.head_dim + 2 * c.head_dim + DA_M_CWARPS * G * (2 + c.head_dim)) * 4 + 4 * c.head_dim + 64);      // slow attention: q, new k/v, per-warp partials
    upd((size_t)3 * (qd / grid + 2) * e->nsplit * 4);
    upd((size_t)3 * grid * 4 + 256);
    upd((size_t)16896 + (192 + 34) * 8 + 80 * 4 + 64);          // slow head: bins / cut list (or sort buffer) + sampler scratch
    upd((size_t)256 * 8 + (4 * 512 + 2 * 512 + 8) * 4 + 512 * 8 + 64);      // fast heads: scratch + bins / cut list
    a.kmax = kmax; a.lg_rows = (2 * (pairs(c.vocab_size) / grid + 1) + 31) / 16 * 16; a.work_bytes = (int)work;
    a.kv_bytes = c.num_codebooks * 2 * fkd * 2;      // one fast layer's K/V rows; all layers live in a per-CTA global scratch
    const int dim_max = c.dim > c.fast_dim ? c.dim : c.fast_dim;
    const MegaSmem fixed = mega_smem_plan(a.kmax, dim_max, a.lg_rows, a.work_bytes, a.kv_bytes, 0);
    const long long budget = 227 * 1024 - 1024 - (long long)fixed.total;   // 1 KB for the kernel's static shared variables
    long long ring = budget / 1024 * 1024;
    if (ring < (long long)(2 * max_entry) || ring < 32 * 1024) { e->use_mega = false; return 0; }
    a.ring_bytes = (int)ring;
    a.plan = mega_smem_plan(a.kmax, dim_max, a.lg_rows, a.work_bytes, a.kv_bytes, a.ring_bytes);
    e->mega_smem = a.plan.total;
  }
  { const char *v = getenv("DUALAR_KEEP_FRAC"); if (v) { float fr = (float)atof(v); CU(cudaMemcpyToSymbol(g_keep_frac, &fr, sizeof(float))); } }
  { const char *v = getenv("DUALAR_KEEP_MODE"); if (v) { int md = atoi(v); CU(cudaMemcpyToSymbol(g_keep_mode, &md, sizeof(int))); } }
  { const char *v = getenv("DUALAR_POLL_NS"); int ns = v ? atoi(v) : 0; CU(cudaMemcpyToSymbol(g_poll_ns, &ns, sizeof(int))); }
  CU(cudaFuncSetAttribute(mega_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->mega_smem));
  CU(cudaFuncSetAttribute(mega_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->mega_smem));
  CU(cudaFuncSetAttribute(mega_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->mega_smem));
  CU(cudaFuncSetAttribute(mega_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->mega_smem));
  return 0;
}

extern "C" int dualar_finalize(dualar_engine *e) {
  if (!e) return fail(DUALAR_EINVAL, "null engine");
  if (e->finalized) return 0;
  const dualar_config &c = e->c;
  CU(cudaSetDevice(e->device));
  if (!e->rope_loaded) {
    std::vector<uint16_t> t; host_rope(t, c.max_seq_len, c.head_dim, c.rope_base);
    CU(cudaMemcpy(e->rope, t.data(), t.size() * 2, cudaMemcpyHostToDevice)); e->slots["freqs_cis"].loaded = true;
  }
  if (!e->fast_rope_loaded) {
    std::vector<uint16_t> t; host_rope(t, c.num_codebooks, c.fast_head_dim, c.rope_base);
    CU(cudaMemcpy(e->fast_rope, t.data(), t.size() * 2, cudaMemcpyHostToDevice)); e->slots["fast_freqs_cis"].loaded = true;
  }
  for (auto &kv : e->slots) if (!kv.second.loaded) return fail(DUALAR_EMISSING, "weight '%s' was never loaded", kv.first.c_str());
  // KV caches (llama.py:126-149 layout) unless the caller bound its own
  for (auto &L : e->slow) if (!L.kv_external) {
    size_t n = (size_t)c.n_local_heads * c.max_seq_len * c.head_dim; int rc;
    if ((rc = dev_alloc(e, L.kc, n))) return rc; if ((rc = dev_alloc(e, L.vc, n))) return rc;
  }
  for (auto &L : e->fast) if (!L.kv_external) {
    size_t n = (size_t)c.fast_n_local_heads * c.num_codebooks * c.fast_head_dim; int rc;
    if ((rc = dev_alloc(e, L.kc, n))) return rc; if ((rc = dev_alloc(e, L.vc, n))) return rc;
  }
  int rc;
  const int G = c.n_head / c.n_local_heads;
  e->nsplit = e->sms / c.n_local_heads; if (e->nsplit < 1) e->nsplit = 1; if (e->nsplit > 64) e->nsplit = 64;
  const int qkv_rows = (c.n_head + 2 * c.n_local_heads) * c.head_dim, fqkv_rows = (c.fast_n_head + 2 * c.fast_n_local_heads) * c.fast_head_dim;
  if ((rc = dev_alloc(e, e->x, c.dim)) || (rc = dev_alloc(e, e->h, c.dim)) || (rc = dev_alloc(e, e->qkv, qkv_rows)) ||
      (rc = dev_alloc(e, e->y, c.n_head * c.head_dim)) || (rc = dev_alloc(e, e->act, c.intermediate_size)) ||
      (rc = dev_alloc(e, e->logits, c.vocab_size)) || (rc = dev_alloc(e, e->logits_raw, c.vocab_size)) ||
      (rc = dev_alloc(e, e->fin, c.fast_dim)) || (rc = dev_alloc(e, e->fpi, c.fast_dim)) || (rc = dev_alloc(e, e->fbuf[0], c.fast_dim)) || (rc = dev_alloc(e, e->fbuf[1], c.fast_dim)) ||
      (rc = dev_alloc(e, e->fh, c.fast_dim)) || (rc = dev_alloc(e, e->fqkv, fqkv_rows)) || (rc = dev_alloc(e, e->fact, c.fast_intermediate_size)) ||
      (rc = dev_alloc(e, e->flogits, (size_t)c.num_codebooks * e->fv)) || (rc = dev_alloc(e, e->flogits_raw, (size_t)c.num_codebooks * e->fv)) ||
      (rc = dev_alloc(e, e->part_o, (size_t)c.n_local_heads * e->nsplit * G * c.head_dim)) ||
      (rc = dev_alloc(e, e->part_ml, (size_t)c.n_local_heads * e->nsplit * G * 2)) ||
      (rc = dev_alloc(e, e->partials, (size_t)2 * e->sms)) || (rc = dev_alloc(e, e->cand, (size_t)DA_CAND_CAP)) ||
      (rc = dev_alloc(e, e->st, 1)) || (rc = dev_alloc(e, e->seq, (size_t)(c.num_codebooks + 1) * c.max_seq_len)))
    return rc;
  CU(cudaMallocHost((void **)&e->h_seq, (size_t)(c.num_codebooks + 1) * c.max_seq_len * sizeof(int)));
  CU(cudaMallocHost((void **)&e->h_st, sizeof(DAState)));
  CU(cudaHostAlloc((void **)&e->h_err_sticky, sizeof(int), cudaHostAllocMapped)); *e->h_err_sticky = 0;
  CU(cudaHostGetDevicePointer((void **)&e->d_err_sticky, e->h_err_sticky, 0));
  CU(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
  { const char *v = getenv("DUALAR_PDL"); if (v && v[0] == '0') g_use_pdl = false; }
  { const char *v = getenv("DUALAR_MEGA"); if (v) e->use_mega = v[0] != '0'; }
  if (e->use_mega && (rc = build_mega(e)) < 0) return rc;
  { const char *v = getenv("DUALAR_TIMELINE"); if (v && v[0] == '1') { e->tl_slots = 2048; if ((rc = dev_alloc(e, e->tl, (size_t)e->tl_slots * 8))) return rc;
      if ((rc = dev_alloc(e, e->tl2, (size_t)DA_M_MAX_PHASES * 160 * 4))) return rc; } }
  // the fast stack is re-read num_codebooks times per token.  Experiment switch: a persisting-L2 carve-out + access-policy window over
  // it (DUALAR_L2_WINDOW=1, hit ratio DUALAR_L2_HIT) instead of the per-copy fractional evict_last hints
  { const char *v = getenv("DUALAR_L2_WINDOW"); e->l2_window = v && v[0] == '1';
    if (e->l2_window) {
      int maxp = 0, l2 = 0, maxw = 0; cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, e->device);
      cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, e->device); cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, e->device);
      if (maxp > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp); cudaGetLastError();
      const char *h = getenv("DUALAR_L2_HIT");
      e->l2_hit_ratio = h ? (float)atof(h) : 0.0f;      // 0: carve-out / window size
      if (e->l2_hit_ratio > 1.0f) e->l2_hit_ratio = 1.0f;
      e->l2_persist_bytes = maxp > 0 ? (size_t)maxp : 0;
      fprintf(stderr, "[dualar] L2 %d MB, max persisting %d MB, max window %d MB, fast stack %.1f MB, hit ratio %s\n", l2 >> 20, maxp >> 20, maxw >> 20,
              e->fast_bytes / 1048576.0, h ? h : "carve-out / window");
      int one = 1; CU(cudaMemcpyToSymbol(g_l2_window, &one, sizeof(int)));
    } }
  // a benign state for the dry runs inside capture()
  DAState init; memset(&init, 0, sizeof(init)); init.temperature = 0.7f; init.top_p = 0.8f; init.rep_penalty = 1.1f;
  init.tok_in[0] = c.semantic_begin_id; init.seed = e->seed; init.cpu_sem = e->cpu_sem;
  CU(cudaMemcpy(e->st, &init, sizeof(init), cudaMemcpyHostToDevice));
  if ((rc = capture(e, false, &e->g_step, &e->launches_step))) return rc;
  if ((rc = capture(e, true, &e->g_prefill, &e->launches_prefill))) return rc;
  CU(cudaMemcpy(e->st, &init, sizeof(init), cudaMemcpyHostToDevice));
  if (e->u_arena) CU(cudaMemset(e->u_arena, 0, e->u_arena_bytes));
  // the dry runs wrote KV position 0 / 1 of caches we own; clear them again
  for (auto &L : e->slow) if (!L.kv_external) {
    size_t n = (size_t)c.n_local_heads * c.max_seq_len * c.head_dim * sizeof(bf16);
    CU(cudaMemset(L.kc, 0, n)); CU(cudaMemset(L.vc, 0, n));
  }
  CU(cudaDeviceSynchronize());
  e->finalized = true;
  return 0;
}

static void batch_destroy(dualar_engine *e);
static int batch_select_path(dualar_engine *e);
static int tc_prefill_own(dualar_engine *e, int t0, int t1, cudaStream_t s);

extern "C" void dualar_destroy(dualar_engine *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  batch_destroy(e);
  if (e->g_step) cudaGraphExecDestroy(e->g_step);
  if (e->g_prefill) cudaGraphExecDestroy(e->g_prefill);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  for (void *p : e->owned) cudaFree(p);
  if (e->h_seq) cudaFreeHost(e->h_seq);
  if (e->h_st) cudaFreeHost(e->h_st);
  if (e->h_err_sticky) cudaFreeHost(e->h_err_sticky);
  for (auto &ev : e->ev_ring) if (ev) cudaEventDestroy(ev);
  if (e->arena) cudaFree(e->arena);
  delete e->ma_step; delete e->ma_prefill;
  delete e;
}

extern "C" int dualar_seed(dualar_engine *e, uint64_t seed) {
  if (!e) return fail(DUALAR_EINVAL, "null engine");
  e->seed = seed;
  if (e->finalized) {
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpy(&e->st->seed, &seed, sizeof(seed), cudaMemcpyHostToDevice));
    unsigned zero = 0; CU(cudaMemcpy(&e->st->step_ctr, &zero, sizeof(zero), cudaMemcpyHostToDevice));
  }
  return 0;
}

extern "C" int dualar_set_option(dualar_engine *e, const char *name, double value) {
  if (!e || !name) return fail(DUALAR_EINVAL, "null argument");
  if (!strcmp(name, "cpu_scalar_semantics")) {
    e->cpu_sem = value != 0.0;
    if (e->finalized) { CU(cudaSetDevice(e->device)); CU(cudaMemcpy(&e->st->cpu_sem, &e->cpu_sem, sizeof(int), cudaMemcpyHostToDevice)); }
    return 0;
  }
  if (!strcmp(name, "prefill_mode")) { e->prefill_mode = value != 0.0; e->kv_dirty = true; return 0; }
  if (!strcmp(name, "prefix_reuse")) { e->prefix_reuse = value != 0.0; return 0; }
  if (!strcmp(name, "batch_group_slots")) {
    if (e->batch) return fail(DUALAR_ESTATE, "batch_group_slots must be set before dualar_batch_init");
    if (value < 1 || value > 128) return fail(DUALAR_EINVAL, "batch_group_slots must be in [1, 128]");
    e->batch_group_slots = (int)value; return 0;
  }
  if (!strcmp(name, "batch_decode_join")) { e->batch_decode_join = value != 0.0; return 0; }
  if (!strcmp(name, "batch_persistent")) { e->batch_persistent = value != 0.0; return batch_select_path(e); }
  if (!strcmp(name, "mega_kernel")) {
    if (e->finalized) return fail(DUALAR_ESTATE, "mega_kernel must be set before dualar_finalize");
    e->use_mega = value != 0.0; return 0;
  }
  if (!strcmp(name, "fast_qkv_table")) {
    if (e->finalized) return fail(DUALAR_ESTATE, "fast_qkv_table must be set before dualar_finalize");
    e->use_t0 = value != 0.0; return 0;
  }
  if (!strcmp(name, "candidate_delta")) {
    if (e->finalized) return fail(DUALAR_ESTATE, "candidate_delta must be set before dualar_finalize");
    if (!(value > 0.0)) return fail(DUALAR_EINVAL, "candidate_delta must be positive");
    e->delta = (float)value; return 0;
  }
  return fail(DUALAR_EINVAL, "unknown option '%s'", name);
}

extern "C" int dualar_set_noise(dualar_engine *e, const void *noise, int64_t n_steps) {
  if (!e) return fail(DUALAR_EINVAL, "null engine");
  e->noise = (const bf16 *)noise; e->noise_steps = noise ? n_steps : 0;
  return 0;
}

extern "C" int dualar_fill_noise(dualar_engine *e, uint64_t seed, uint32_t step, uint32_t head, void *out, int64_t n, void *stream) {
  if (!e || !out || n < 0) return fail(DUALAR_EINVAL, "bad argument");
  CU(cudaSetDevice(e->device));
  if (n == 0) return 0;
  int grid = (int)((n + 255) / 256); if (grid > 4 * e->sms) grid = 4 * e->sms;
  fill_noise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((bf16 *)out, n, seed, step, head);
  CU(cudaGetLastError());
  return 0;
}

// ---- step mode (decode_one_token_ar, inference.py:83-155) --------------------------------------------
extern "C" int dualar_step(dualar_engine *e, const int32_t *x, const int32_t *input_pos, const int32_t *prev, int64_t prev_stride,
                           const float *temperature, const float *top_p, const float *rep, const void *noise, int32_t *out, void *stream) {
  if (!e || !x || !input_pos || !temperature || !top_p || !rep || !out) return fail(DUALAR_EINVAL, "null argument");
  if (!e->finalized) return fail(DUALAR_ESTATE, "dualar_finalize has not been called");
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  LoadStepArgs a{x, input_pos, prev, prev_stride, temperature, top_p, rep, (const bf16 *)noise, e->c.num_codebooks + 1, e->st};
  load_step_kernel<<<1, 256, 0, s>>>(a); CU(cudaGetLastError());
  CU(cudaGraphLaunch(e->g_step, s));
  store_step_kernel<<<1, 32, 0, s>>>(e->st, out, e->c.num_codebooks + 1, e->d_err_sticky); CU(cudaGetLastError());
  e->request_open = false; e->kv_dirty = true;
  // the step is asynchronous, like the reference's compiled step: a device fault raised by an EARLIER step (the kernels record it
  // in a sticky word that store_step_kernel copies to mapped host memory) is reported by the next call
  if (e->h_err_sticky && *e->h_err_sticky) {
    const int code = *e->h_err_sticky; *e->h_err_sticky = 0;
    return fail(DUALAR_EDEVICE, "device fault flag %d in an earlier step (1: token id out of range, 2: bulk-copy wait timed out, 3: code >= codebook_size, 4: hand-over poll timed out)", code);
  }
  return 0;
}

// ---- test hook: the sampler alone on caller-supplied logits ------------------------------------------------
namespace da {
__global__ void __launch_bounds__(DA_GEMV_THREADS, 1) debug_fast_sample_kernel(const GemvArgs a, const bf16 *in) {
  extern __shared__ __align__(16) float smem_dbg[];
  DAState *st = a.st;
  const float rp_bf = eff_rep_penalty(st);
  for (int i = threadIdx.x; i < a.rows; i += blockDim.x) {
    float z = bf2f(in[i]);
    if (st->use_penalty) for (int c = 0; c < DA_WIN; ++c) if (st->win[(a.head + 1) * DA_WIN + c] == i) { z = penalise(z, rp_bf); break; }
    a.out[i] = f2bf(z);
  }
  __threadfence();
  __syncthreads();
  fast_head_sample(a, smem_dbg);
}
}  // namespace da

extern "C" int dualar_debug_sample(dualar_engine *e, int head, const void *logits, const int32_t *prev, int64_t prev_stride,
                                   const float *temperature, const float *top_p, const float *rep, const void *noise,
                                   int32_t *out_token, void *stream) {
  if (!e || !logits || !temperature || !top_p || !rep || !out_token) return fail(DUALAR_EINVAL, "null argument");
  if (!e->finalized) return fail(DUALAR_ESTATE, "dualar_finalize has not been called");
  const dualar_config &c = e->c;
  if (head < 0 || head >= c.num_codebooks) return fail(DUALAR_EINVAL, "head out of range");
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  LoadStepArgs la{e->st->tok_in, &e->st->pos, prev, prev_stride, temperature, top_p, rep, (const bf16 *)noise, c.num_codebooks + 1, e->st};
  load_step_kernel<<<1, 256, 0, s>>>(la); CU(cudaGetLastError());
  e->kv_dirty = true;
  if (e->use_mega) {
    // the PRODUCT path's sampler: one whole decode step of the persistent kernel whose logits epilogues are fed the caller's logits
    // (MegaArgs::force_slow / force_fast), so the repetition penalty, the per-CTA statistics, the ordered candidate list, the binned /
    // sorting / whole-vocabulary samplers and the fast-head samplers that run in production are the ones under test
    MegaArgs a = *e->ma_step;
    a.tl = nullptr; a.tl_slots = 0; a.tl2 = nullptr;
    bf16 *ff = nullptr;
    if (head == 0) a.force_slow = (const bf16 *)logits;
    else {
      // only head `head`'s slice is the caller's; the other fast heads sample from zeros (their ids are not read)
      if (!e->dbg_fast) { int rc = dev_alloc(e, e->dbg_fast, (size_t)(c.num_codebooks - 1) * e->fv); if (rc) return rc; }
      ff = e->dbg_fast;
      CU(cudaMemsetAsync(ff, 0, (size_t)(c.num_codebooks - 1) * e->fv * sizeof(bf16), s));
      CU(cudaMemcpyAsync(ff + (size_t)(head - 1) * e->fv, logits, (size_t)e->fv * sizeof(bf16), cudaMemcpyDeviceToDevice, s));
      a.force_fast = ff;
    }
    if (e->u_arena) CU(cudaMemsetAsync(e->u_arena, 0, e->u_arena_bytes, s));
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(e->sms); cfg.blockDim = dim3(DA_M_THREADS); cfg.dynamicSmemBytes = e->mega_smem; cfg.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CU(cudaLaunchKernelEx(&cfg, mega_kernel<false, true>, a));
    CU(cudaMemcpyAsync(out_token, &e->st->tok_out[head == 0 ? 0 : head + 1], sizeof(int), cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  if (head == 0) {
    debug_stats_kernel<<<e->sms, 512, 0, s>>>((const bf16 *)logits, e->logits, e->partials, c.vocab_size, c.num_codebooks + 1, e->st); CU(cudaGetLastError());
    SelectArgs a; memset(&a, 0, sizeof(a));
    a.logits = e->logits; a.partials = e->partials; a.n_partials = e->sms; a.V = c.vocab_size; a.delta = e->delta; a.cand = e->cand;
    a.fast_emb = e->fast_emb; a.fast_x = e->fin; a.fast_dim = c.fast_dim; a.codebook_size = c.codebook_size; a.sem_begin = c.semantic_begin_id; a.st = e->st;
    size_t smem = 192 * 8 + 34 * 8 + 80 * 4 + 64;
    select_sample_kernel<<<e->sms, 512, smem, s>>>(a); CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_token, &e->st->tok_out[0], sizeof(int), cudaMemcpyDeviceToDevice, s));
  } else {
    GemvArgs a = base_args(e->fast_out, nullptr, e->fv, c.fast_dim, 1);
    a.out = e->flogits + (size_t)(head - 1) * e->fv; a.head = head; a.noise_off = (long long)c.vocab_size + (long long)(head - 1) * e->fv;
    a.fast_emb = e->fast_emb; a.fast_x = e->fin; a.fast_dim = c.fast_dim; a.codebook_size = c.codebook_size; a.last_head = 0;
    a.n_rows_tok = c.num_codebooks + 1; a.st = e->st;
    size_t smem = 192 * 8 + 80 * 4 + 64;
    debug_fast_sample_kernel<<<1, DA_GEMV_THREADS, smem, s>>>(a, (const bf16 *)logits); CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_token, &e->st->tok_out[head + 1], sizeof(int), cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

// ---- loop mode (generate / generate_streaming, inference.py:279-384, 643-738) ---------------------------
extern "C" int dualar_prefill(dualar_engine *e, const int32_t *prompt, int T, int max_new, float temperature, float top_p, float rep, void *stream) {
  if (!e || !prompt) return fail(DUALAR_EINVAL, "null argument");
  if (!e->finalized) return fail(DUALAR_ESTATE, "dualar_finalize has not been called");
  const dualar_config &c = e->c;
  const int R = c.num_codebooks + 1;
  if (T < 1) return fail(DUALAR_EINVAL, "empty prompt");
  if (T >= c.max_seq_len) return fail(DUALAR_EINVAL, "Input sequence length %d exceeds max_seq_len %d", T, c.max_seq_len);   // inference.py:296-299
  if (max_new <= 0 || T + max_new > c.max_seq_len) max_new = c.max_seq_len - T;                                              // inference.py:301-307
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  CU(cudaStreamSynchronize(s));   // h_seq / h_st staging buffers are reused
  for (int r = 0; r < R; ++r) memcpy(e->h_seq + (size_t)r * T, prompt + (size_t)r * T, (size_t)T * sizeof(int));
  CU(cudaMemcpy2DAsync(e->seq, (size_t)c.max_seq_len * sizeof(int), e->h_seq, (size_t)T * sizeof(int), (size_t)T * sizeof(int), R, cudaMemcpyHostToDevice, s));
  DAState *h = e->h_st; memset(h, 0, sizeof(*h));
  const bool tc_pf = e->prefill_mode == 0 && T > 1;
  // KV reuse: the reference re-runs the whole prompt for every utterance although the VoiceProfile part of it
  // never changes (synthesizer.py:363-377, 415-429).  The KV rows of the positions this prompt SHARES with the previous
  // request's prompt are still in the cache -- decode only ever writes rows >= the prompt length -- so the tensor-core prefill
  // starts at the first differing position.
  int reuse = 0;
  if (tc_pf && e->prefix_reuse && !e->kv_dirty) {
    const int lim = std::min(T - 1, e->prev_T);      // the last prompt position always goes through the decode step
    while (reuse < lim) {
      bool same = true;
      for (int r = 0; r < R && same; ++r) same = prompt[(size_t)r * T + reuse] == e->prev_prompt[(size_t)r * e->prev_T + reuse];
      if (!same) break;
      ++reuse;
    }
  }
  e->prev_prompt.assign(prompt, prompt + (size_t)R * T); e->prev_T = T; e->kv_dirty = !tc_pf; e->last_reuse = reuse;
  h->pos = tc_pf ? T - 1 : 0; h->n_gen = 0; h->max_gen = max_new; h->prompt_len = T; h->loop_mode = 1; h->use_penalty = 0;
  h->temperature = temperature; h->top_p = top_p; h->rep_penalty = rep; h->seed = e->seed; h->step_ctr = 0;
  h->cpu_sem = e->cpu_sem;
  h->noise = e->noise; h->noise_stride = (long long)c.vocab_size + (long long)(c.num_codebooks - 1) * e->fv;
  for (int r = 0; r < R; ++r) h->tok_in[r] = prompt[(size_t)r * T + (tc_pf ? T - 1 : 0)];
  CU(cudaMemcpyAsync(e->st, h, sizeof(*h), cudaMemcpyHostToDevice, s));
  if (e->u_arena) CU(cudaMemsetAsync(e->u_arena, 0, e->u_arena_bytes, s));      // no unit of an earlier request carries a valid tag
  if (tc_pf) {
    // the reference prefills the whole prompt in ONE forward (inference.py:353-362); here positions [0, T-1) go through the
    // tensor-core GEMMs (KV rows only, no head), the last position through the decode step below, which samples the first token
    int rc = tc_prefill_own(e, reuse, T - 1, s); if (rc) return rc;
  } else for (int t = 0; t + 1 < T; ++t) CU(cudaGraphLaunch(e->g_prefill, s));
  CU(cudaGraphLaunch(e->g_step, s));
  e->prompt_len = T; e->max_gen = max_new; e->request_open = true; e->stream_cols = 0; e->steps_enqueued = 0;
  return 0;
}

extern "C" int dualar_decode(dualar_engine *e, int n_steps, void *stream) {
  if (!e) return fail(DUALAR_EINVAL, "null engine");
  if (!e->request_open) return fail(DUALAR_ESTATE, "dualar_decode without a prefilled request");
  CU(cudaSetDevice(e->device));
  for (int i = 0; i < n_steps; ++i) CU(cudaGraphLaunch(e->g_step, (cudaStream_t)stream));
  e->steps_enqueued += n_steps;
  return 0;
}

// ---- streaming hand-off (generate_streaming, inference.py:643-738; synthesize_stream, synthesizer.py:483-584) ---------------------
// n_steps decode steps, then -- on the same stream, so without any host wait in between -- an asynchronous copy of the columns
// they produced into a caller-owned (pinned) HOST buffer and of (n_gen, done) into host_state, then an event.  The host can enqueue
// the NEXT chunk before it waits for this one, so the decode stream never idles while the codec consumes a chunk.
extern "C" int dualar_decode_async(dualar_engine *e, int n_steps, int32_t *host_out, int32_t *host_state, void *stream, int *ticket) {
  if (!e || !host_out || !host_state || !ticket) return fail(DUALAR_EINVAL, "null argument");
  if (!e->request_open) return fail(DUALAR_ESTATE, "dualar_decode_async without a prefilled request");
  if (n_steps < 0) return fail(DUALAR_EINVAL, "n_steps < 0");
  const dualar_config &c = e->c; const int R = c.num_codebooks + 1;
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  // column j of the request lives at seq[:, prompt_len + j]; the prefill call itself produced column 0.  Columns handed out so
  // far: stream_cols; columns that exist once these steps have run: at most 1 + steps enqueued (fewer after <|im_end|>: the host
  // reads n_gen from host_state to know how many of the copied columns are real)
  const int first = e->stream_cols;
  int avail = 1 + e->steps_enqueued + n_steps; if (avail > e->max_gen) avail = e->max_gen;
  int n = avail - first; if (n < 0) n = 0; if (n > n_steps + 1) n = n_steps + 1;
  for (int i = 0; i < n_steps; ++i) CU(cudaGraphLaunch(e->g_step, s));
  e->steps_enqueued += n_steps;
  if (n > 0) CU(cudaMemcpy2DAsync(host_out, (size_t)(n_steps + 1) * sizeof(int), e->seq + e->prompt_len + first, (size_t)c.max_seq_len * sizeof(int),
                                  (size_t)n * sizeof(int), R, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(host_state, &e->st->n_gen, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(host_state + 1, &e->st->done, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(host_state + 2, &e->st->err, sizeof(int), cudaMemcpyDeviceToHost, s));
  host_state[3] = first; host_state[4] = n;
  e->stream_cols = first + n;
  const int t = e->ev_next; e->ev_next = (e->ev_next + 1) % 8;
  if (!e->ev_ring[t]) CU(cudaEventCreateWithFlags(&e->ev_ring[t], cudaEventDisableTiming));
  CU(cudaEventRecord(e->ev_ring[t], s));
  *ticket = t;
  return 0;
}
// block = 0: returns 0 when the chunk has landed, 1 when it has not; block = 1: waits for it
extern "C" int dualar_wait(dualar_engine *e, int ticket, int block) {
  if (!e || ticket < 0 || ticket >= 8 || !e->ev_ring[ticket]) return fail(DUALAR_EINVAL, "bad ticket");
  if (block) { CU(cudaEventSynchronize(e->ev_ring[ticket])); return 0; }
  cudaError_t q = cudaEventQuery(e->ev_ring[ticket]);
  if (q == cudaSuccess) return 0;
  if (q == cudaErrorNotReady) return 1;
  return fail(DUALAR_ECUDA, "cudaEventQuery: %s", cudaGetErrorString(q));
}

extern "C" int dualar_collect(dualar_engine *e, int32_t *out, int cap, int *n_tokens, int *finished, void *stream) {
  if (!e || !n_tokens) return fail(DUALAR_EINVAL, "null argument");
  if (!e->request_open) return fail(DUALAR_ESTATE, "dualar_collect without a prefilled request");
  const dualar_config &c = e->c; const int R = c.num_codebooks + 1;
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  CU(cudaMemcpyAsync(e->h_st, e->st, sizeof(DAState), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (e->h_st->err) return fail(DUALAR_EDEVICE, "device fault flag %d (1: token id out of range, 2: bulk-copy wait timed out, 3: code >= codebook_size, 4: hand-over poll timed out)", e->h_st->err);
  int n = e->h_st->n_gen;
  *n_tokens = n;
  if (finished) *finished = e->h_st->done;
  if (out && n > 0) {
    if (cap < n) return fail(DUALAR_EINVAL, "output capacity %d < %d generated columns", cap, n);
    CU(cudaMemcpy2DAsync(e->h_seq, (size_t)n * sizeof(int), e->seq + e->prompt_len, (size_t)c.max_seq_len * sizeof(int), (size_t)n * sizeof(int), R, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (int r = 0; r < R; ++r) memcpy(out + (size_t)r * cap, e->h_seq + (size_t)r * n, (size_t)n * sizeof(int));
  }
  return 0;
}

extern "C" int dualar_generate(dualar_engine *e, const int32_t *prompt, int T, int max_new, float temperature, float top_p, float rep,
                               int32_t *out, int cap, int *n_tokens, void *stream) {
  int rc = dualar_prefill(e, prompt, T, max_new, temperature, top_p, rep, stream);
  if (rc) return rc;
  // every replay after <|im_end|> is a no-op, so the whole budget can be enqueued without a host round trip;
  // checking in on the device flag every 64 steps keeps the wasted launches bounded
  int remaining = e->max_gen - 1, n = 0, fin = 0;
  while (remaining > 0) {
    int chunk = remaining < 64 ? remaining : 64;
    if ((rc = dualar_decode(e, chunk, stream))) return rc;
    remaining -= chunk;
    if (remaining > 0) {
      CU(cudaMemcpyAsync(e->h_st, e->st, sizeof(DAState), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
      CU(cudaStreamSynchronize((cudaStream_t)stream));
      if (e->h_st->done) break;
    }
  }
  rc = dualar_collect(e, out, cap, &n, &fin, stream);
  if (n_tokens) *n_tokens = n;
  return rc;
}

extern "C" int dualar_read_buffer(dualar_engine *e, const char *name, void *dst, int64_t nbytes, void *stream) {
  if (!e || !name || !dst) return fail(DUALAR_EINVAL, "null argument");
  if (!e->finalized) return fail(DUALAR_ESTATE, "dualar_finalize has not been called");
  const dualar_config &c = e->c;
  const void *src = nullptr; int64_t avail = 0;
  if (!strcmp(name, "slow_logits")) { src = e->logits; avail = (int64_t)c.vocab_size * 2; }
  else if (!strcmp(name, "slow_logits_raw")) { src = e->logits_raw; avail = (int64_t)c.vocab_size * 2; }
  else if (!strcmp(name, "hidden")) { src = e->x; avail = (int64_t)c.dim * 2; }
  else if (!strcmp(name, "fast_logits")) { src = e->flogits_raw; avail = (int64_t)(c.num_codebooks - 1) * e->fv * 2; }
  else if (!strcmp(name, "tokens")) { src = e->st->tok_out; avail = (int64_t)(c.num_codebooks + 1) * 4; }
  else if (!strcmp(name, "nucleus")) { src = e->st->nucleus; avail = (int64_t)c.num_codebooks * 4; }
  else if (!strcmp(name, "n_cand")) { src = &e->st->n_cand; avail = 4; }
  else if (!strcmp(name, "prefix_reused") || !strcmp(name, "prefill_launches")) {      // host-side counters of the last dualar_prefill
    if (nbytes < 4) return fail(DUALAR_EINVAL, "4 bytes needed");
    *(int *)dst = !strcmp(name, "prefix_reused") ? e->last_reuse : e->prefill_launches; return 0;
  }
  else if (!strcmp(name, "qkv")) { src = e->qkv; avail = (int64_t)(c.n_head + 2 * c.n_local_heads) * c.head_dim * 2; }
  else if (!strcmp(name, "y")) { src = e->y; avail = (int64_t)c.n_head * c.head_dim * 2; }
  else if (!strcmp(name, "h")) { src = e->h; avail = (int64_t)c.dim * 2; }
  else if (!strcmp(name, "act")) { src = e->act; avail = (int64_t)c.intermediate_size * 2; }
  else if (!strcmp(name, "fast_x")) { src = e->fbuf[(c.n_fast_layer - 1) & 1]; avail = (int64_t)c.fast_dim * 2; }
  else if (!strcmp(name, "timeline")) { if (!e->tl) return fail(DUALAR_ESTATE, "run with DUALAR_TIMELINE=1"); src = e->tl; avail = (int64_t)e->tl_slots * 64; }
  else if (!strcmp(name, "fast_in")) { src = e->fin; avail = (int64_t)c.fast_dim * 2; }
  else if (!strcmp(name, "timeline2")) { if (!e->tl2) return fail(DUALAR_ESTATE, "run with DUALAR_TIMELINE=1"); src = e->tl2; avail = (int64_t)DA_M_MAX_PHASES * 160 * 4 * 8; }
  else if (!strcmp(name, "cand")) { if (!e->m_cand) return fail(DUALAR_ESTATE, "persistent kernel not in use"); src = e->m_cand; avail = (int64_t)DA_CAND_CAP * 8; }
  else return fail(DUALAR_EINVAL, "unknown buffer '%s'", name);
  if (nbytes > avail) return fail(DUALAR_EINVAL, "buffer '%s' holds %lld bytes, %lld requested", name, (long long)avail, (long long)nbytes);
  CU(cudaSetDevice(e->device));
  // under the persistent kernel the activation vectors live as tagged 32-bit units (bf16 value in the high half)
  const uint32_t *units = nullptr;
  if (e->use_mega) {
    if (!strcmp(name, "hidden")) units = e->u_x;
    else if (!strcmp(name, "qkv")) units = e->u_qkv;
    else if (!strcmp(name, "y")) units = e->u_y;
    else if (!strcmp(name, "h")) units = e->u_h;
    else if (!strcmp(name, "act")) units = e->u_act;
    else if (!strcmp(name, "fast_x")) units = (c.n_fast_layer - 1) & 1 ? e->u_fx1 : e->u_fx0;
    else if (!strcmp(name, "fast_in")) units = e->u_fin;
  }
  if (units) {
    const size_t n = (size_t)nbytes / 2;
    std::vector<uint32_t> tmp(n);
    CU(cudaMemcpyAsync(tmp.data(), units, n * 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    uint16_t *o = (uint16_t *)dst;
    for (size_t i = 0; i < n; ++i) o[i] = (uint16_t)(tmp[i] >> 16);
    return 0;
  }
  CU(cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

extern "C" int dualar_launches_per_step(const dualar_engine *e, int *decode_step, int *prefill_position) {
  if (!e || !e->finalized) return fail(DUALAR_ESTATE, "engine not finalized");
  if (decode_step) *decode_step = e->launches_step;
  if (prefill_position) *prefill_position = e->launches_prefill;
  return 0;
}

extern "C" int dualar_weight_bytes(const dualar_engine *e, int64_t *total, int64_t *fast) {
  if (!e) return fail(DUALAR_EINVAL, "null engine");
  if (total) *total = (int64_t)e->arena_bytes;
  if (fast) *fast = (int64_t)e->fast_bytes;
  return 0;
}

#include "batch_host.cuh"

static int tc_prefill_own(dualar_engine *e, int t0, int t1, cudaStream_t s) {
  const dualar_config &c = e->c;
  std::vector<bf16 *> kc(c.n_layer), vc(c.n_layer);
  for (int l = 0; l < c.n_layer; ++l) { kc[l] = e->slow[l].kc; vc[l] = e->slow[l].vc; }
  KvTarget kv{kc.data(), vc.data(), 0, c.max_seq_len};
  return tc_prefill(e, kv, e->seq, c.max_seq_len, t0, t1, s);      // (kernel arguments are copied at launch: the vectors may go out of scope)
}
