// batch_host.cuh -- host side of the many-column path: the tensor-core prefill of a prompt and batched decode.
// Included at the end of engine.cu (it needs dualar_engine).  See include/dualar.h for the contract of the entry points.
//
// One schedule serves both uses: a COLUMN is a prompt position (prefill: T columns of one request, KV rows written for all of
// them, no LM head) or a request slot (batched decode: B columns, one per slot, each with its own KV cache, position, sampling
// state and Philox stream).  Per slow layer: RMSNorm -> wqkv GEMM -> q/k norm + RoPE + KV write (a kernel of its own in prefill,
// inside the attention kernel in decode) -> split-KV attention -> wo GEMM (+ residual) -> RMSNorm -> w1/w3 GEMM (SwiGLU epilogue)
// -> w2 GEMM (+ residual); all GEMMs are gemm_tc_kernel (tcgen05), their K splits a thread-block cluster.
// Request slots live in GROUPS (one graph, stream and set of buffers each; the groups of a step run concurrently), prefills run on
// the engine's own stream beside them; see dualar_batch_init.  Experiment switches (environment, read once per engine):
//   DUALAR_TC_CLUSTER=0        split-K through a global workspace + ticket instead of the cluster reduction
//   DUALAR_TC_CTAS / _MINKB    split heuristic: CTAs aimed at per GEMM, least k-blocks per CTA;  DUALAR_TC_STAGES: ring depth
//   DUALAR_TC_FUSE_NORM=1 / 2  RMSNorm inside the consuming GEMM (own statistics / statistics from the producing GEMM) -- slower
//   DUALAR_ATTN_MMA=0          scalar fp32 attention walk instead of the tensor-core one;  DUALAR_ATTN_FUSE_POST=0: separate q/k-norm kernel
//   DUALAR_ATTN_CLUSTER=0, DUALAR_ATTN_TPS     KV-split merge through a ticket for small engines too; least tiles per split
//   DUALAR_BATCH_GROUP_SLOTS, DUALAR_BATCH_FORK, DUALAR_BATCH_PERSIST, DUALAR_BS_*   group size, LM-head fork, persistent step (bstep.cuh)
#pragma once

namespace {

#define DA_SSQ_LD 64      // row tiles per column a statistics buffer has room for (rows <= 8192)

struct MapKey { const void *p; int rows, K, box; bool operator<(const MapKey &o) const { return std::tie(p, rows, K, box) < std::tie(o.p, o.rows, o.K, o.box); } };

// activation buffers for up to `cap` columns
struct ColBufs {
  int cap = 0;
  bf16 *x = nullptr, *xn = nullptr, *qkv = nullptr, *y = nullptr, *h = nullptr, *act = nullptr;
  float *part_o = nullptr, *part_ml = nullptr; unsigned int *attn_tickets = nullptr;
  int nsplit = 1;
  // decode only
  bf16 *logits = nullptr, *logits_raw = nullptr, *fin = nullptr, *fx[2] = {nullptr, nullptr}, *fh = nullptr, *fqkv = nullptr, *fy = nullptr, *fact = nullptr,
       *fxn = nullptr, *flogits = nullptr, *flogits_raw = nullptr, *fpi = nullptr;
  float *cmax = nullptr; unsigned long long *cand = nullptr;
  // DUALAR_TC_FUSE_NORM=2: per activation buffer that feeds an RMSNorm, the partial sums of squares its producing GEMM leaves ([cap][DA_SSQ_LD]);
  // `ssq_valid` tracks, while a step is being enqueued, which buffers were last written by such a GEMM
  std::map<const bf16 *, float *> ssq; std::set<const bf16 *> ssq_valid;
};

// the step's kernel sequence recorded as phases of the persistent kernel (bstep.cuh) instead of being launched
struct BRec {
  std::vector<BPhase> ph; std::vector<CUtensorMap> maps; std::map<const CUtensorMap *, int> idx;
  int split = -1;      // phases [0, split) run before the slow sampler's kernels, [split, n) after them
  BPhase &add(int kind, dim3 g) { BPhase p; memset(&p, 0, sizeof(p)); p.kind = kind; p.gx = (int)g.x; p.gy = (int)g.y; p.gz = (int)g.z; ph.push_back(p); return ph.back(); }
  int map_index(const CUtensorMap *m) { auto it = idx.find(m); if (it != idx.end()) return it->second; const int i = (int)maps.size(); maps.push_back(*m); idx[m] = i; return i; }
};

}  // namespace

struct dualar_batch {
  int B = 0, BN = 32, Sb = 0, nchunk = 32;
  ColBufs c;
  DAState *st = nullptr, *h_st = nullptr; int *seq = nullptr, *h_seq = nullptr;      // h_st / h_seq: pinned staging of dualar_batch_collect
  DAState *h_act = nullptr; int *h_prompt = nullptr;      // pinned staging of dualar_batch_prefill: one state per slot, one prompt buffer
  cudaEvent_t ev_prompt = nullptr; bool prompt_pending = false, decode_pending = false; std::vector<char> known_done;
  std::vector<bf16 *> kc, vc, fkc, fvc; long long slot_stride = 0, fslot_stride = 0;
  cudaGraphExec_t g_step = nullptr, g_step_p = nullptr; int launches = 0, launches_p = 0;      // per-kernel graph / persistent graph
  cudaStream_t side_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;      // LM head + slow sampler run beside fast pass 0
  std::vector<int> prompt_len, max_gen; std::vector<char> open;
  float *ws = nullptr; unsigned int *tickets = nullptr;      // split-K workspace of THIS group's GEMMs (groups run concurrently)
  cudaStream_t stream = nullptr; cudaEvent_t ev_done = nullptr;      // the group's own stream when there are several groups
  cudaEvent_t ev_pf = nullptr, ev_act = nullptr, ev_rel = nullptr; bool act_pending = false, rel_pending = false;      // asynchronous prefill: KV rows written / slot activated
  DAState *st_idle = nullptr;      // device copy of the idle slot state (loop_mode 0, position -1: the step's KV write and attention skip the slot)
  // persistent step (bstep.cuh): two cooperative launches around the slow sampler's kernels
  bool persistent = false; BPhase *d_phases = nullptr; CUtensorMap *d_maps = nullptr; unsigned int *gbar = nullptr;
  int n_phases = 0, split = 0, bs_stages = 0; size_t bs_smem = 0; long long *d_tl = nullptr; std::vector<int> kinds;
};

struct dualar_tc {
  std::map<MapKey, CUtensorMap> maps;
  float *ws = nullptr; size_t ws_bytes = 0; unsigned int *tickets = nullptr; int *err = nullptr;      // ws / tickets: the workspace GEMMs are enqueued with (a group's while its step is captured)
  float *ws_own = nullptr; unsigned int *tickets_own = nullptr;
  ColBufs pf;                 // prefill columns (cap 256)
  bool ready = false;
  int target_ctas = 0, min_kb = 0;      // decode split-K heuristic: CTAs aimed at per GEMM, least k-blocks per CTA (0 = by mode; DUALAR_TC_CTAS, DUALAR_TC_MINKB)
  bool attn_mma = true;            // DUALAR_ATTN_MMA=0: the scalar fp32 walk instead of Q.K^T and P@V on the tensor cores (b_attn_body<B, true>) in decode
  bool attn_cluster = true;        // DUALAR_ATTN_CLUSTER=0: KV splits merged through a global buffer + ticket
  bool cluster_reduce = true;      // DUALAR_TC_CLUSTER=0: split-K through the global workspace + ticket (the prefill path's way)
  bool attn_fuse_post = true;      // DUALAR_ATTN_FUSE_POST=0: b_qkv_post_kernel in front of the decode attention, as in prefill
  bool concurrent_groups = false;      // set while the steps of a multi-group engine are captured
  int attn_tiles_per_split = 0;      // 0 = by mode (see enqueue_slow_cols);      // a KV split is worth its partials / ticket / merge only from this many 64-position tiles on (DUALAR_ATTN_TPS)
  BRec *rec = nullptr;        // non-null while the step is being recorded for the persistent kernel
  int ksplit_override = 0, stages_override = 0, bn_override = 0;
  int fuse_norm = 0;          // DUALAR_TC_FUSE_NORM=1: the decode GEMMs normalise their own operand (gemm_tc_kernel<32, true>) instead of a separate
                              // RMSNorm kernel in front of them.  Bit-identical (tests/test_gpu_batch.py), 145 kernels fewer per step -- and measured
                              // SLOWER on B200 (5.45 vs 4.83 ms per bs-32 step): every CTA re-reads the full rows for the statistics and stages the
                              // operand behind the dependency wait, which costs more than the ~3 us a separate 8-CTA kernel adds to the chain
};

static int tc_map(dualar_engine *e, const void *p, int rows, int K, int box, const CUtensorMap **out) {
  MapKey k{p, rows, K, box};
  auto it = e->tc->maps.find(k);
  if (it == e->tc->maps.end()) {
    CUtensorMap m;
    if (!tc_make_map(&m, p, rows, K, box)) return fail(DUALAR_ECUDA, "cuTensorMapEncodeTiled failed (rows %d, K %d, box %d)", rows, K, box);
    it = e->tc->maps.emplace(k, m).first;
  }
  *out = &it->second;
  return 0;
}

template <int BN> static int tc_configure() {
  int mx = (int)((227 * 1024 - 4096) / (DA_TC_A_BYTES + BN * 128)); if (mx > DA_TC_MAX_STAGES) mx = DA_TC_MAX_STAGES;
  CU(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_tc_smem(BN, mx)));
  return 0;
}
static int tc_max_stages(int BN) { int mx = (int)((227 * 1024 - 4096) / (DA_TC_A_BYTES + BN * 128)); return mx > DA_TC_MAX_STAGES ? DA_TC_MAX_STAGES : mx; }

static int alloc_cols(dualar_engine *e, ColBufs &c, int cap, int nsplit, bool decode) {
  const dualar_config &cf = e->c;
  const int qkv_rows = (cf.n_head + 2 * cf.n_local_heads) * cf.head_dim, qd = cf.n_head * cf.head_dim, G = cf.n_head / cf.n_local_heads;
  const int fqkv_rows = (cf.fast_n_head + 2 * cf.fast_n_local_heads) * cf.fast_head_dim, fqd = cf.fast_n_head * cf.fast_head_dim;
  c.cap = cap; c.nsplit = nsplit;
  int rc;
  if ((rc = dev_alloc(e, c.x, (size_t)cap * cf.dim)) || (rc = dev_alloc(e, c.xn, (size_t)cap * cf.dim)) || (rc = dev_alloc(e, c.qkv, (size_t)cap * qkv_rows)) ||
      (rc = dev_alloc(e, c.y, (size_t)cap * qd)) || (rc = dev_alloc(e, c.h, (size_t)cap * cf.dim)) || (rc = dev_alloc(e, c.act, (size_t)cap * cf.intermediate_size)) ||
      (rc = dev_alloc(e, c.part_o, (size_t)cap * cf.n_local_heads * nsplit * G * cf.head_dim)) || (rc = dev_alloc(e, c.part_ml, (size_t)cap * cf.n_local_heads * nsplit * G * 2)) ||
      (rc = dev_alloc(e, c.attn_tickets, (size_t)cap * cf.n_local_heads)))
    return rc;
  if (decode && e->tc->fuse_norm == 2) {
    for (bf16 *p : {c.x, c.h}) { float *q = nullptr; if ((rc = dev_alloc(e, q, (size_t)cap * DA_SSQ_LD))) return rc; c.ssq[p] = q; }
  }
  if (!decode) return 0;
  if ((rc = dev_alloc(e, c.logits, (size_t)cap * cf.vocab_size)) || (rc = dev_alloc(e, c.logits_raw, (size_t)cap * cf.vocab_size)) ||
      (rc = dev_alloc(e, c.fin, (size_t)cap * cf.fast_dim)) || (rc = dev_alloc(e, c.fx[0], (size_t)cap * cf.fast_dim)) || (rc = dev_alloc(e, c.fx[1], (size_t)cap * cf.fast_dim)) ||
      (rc = dev_alloc(e, c.fh, (size_t)cap * cf.fast_dim)) || (rc = dev_alloc(e, c.fqkv, (size_t)cap * fqkv_rows)) || (rc = dev_alloc(e, c.fy, (size_t)cap * fqd)) ||
      (rc = dev_alloc(e, c.fact, (size_t)cap * cf.fast_intermediate_size)) || (rc = dev_alloc(e, c.fxn, (size_t)cap * cf.fast_dim)) ||
      (rc = dev_alloc(e, c.fpi, (size_t)cap * cf.fast_dim)) ||
      (rc = dev_alloc(e, c.flogits, (size_t)cap * e->fv)) || (rc = dev_alloc(e, c.flogits_raw, (size_t)cap * (cf.num_codebooks - 1) * e->fv)) ||
      (rc = dev_alloc(e, c.cmax, (size_t)cap * 64)) || (rc = dev_alloc(e, c.cand, (size_t)cap * DA_CAND_CAP)))
    return rc;
  if (e->tc->fuse_norm == 2) {
    for (bf16 *p : {c.fx[0], c.fx[1], c.fh, c.fpi}) { float *q = nullptr; if ((rc = dev_alloc(e, q, (size_t)cap * DA_SSQ_LD))) return rc; c.ssq[p] = q; }
  }
  return 0;
}

// one-time set-up of the tensor-core path (tensor-map cache, split-K workspace, the prefill column buffers)
static int tc_init(dualar_engine *e) {
  if (e->tc && e->tc->ready) return 0;
  if (!e->tc) e->tc = new dualar_tc();
  if (!tc_encode_fn()) return fail(DUALAR_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  int rc;
  if ((rc = tc_configure<32>()) || (rc = tc_configure<64>()) || (rc = tc_configure<128>()) || (rc = tc_configure<256>())) return rc;
  CU(cudaFuncSetAttribute(gemm_tc_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CU(cudaFuncSetAttribute(gemm_tc_kernel<32, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_tc_smem_cl(32, tc_max_stages(32) > 10 ? 10 : tc_max_stages(32))));
  CU(cudaFuncSetAttribute(gemm_tc_kernel<64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_tc_smem_cl(64, tc_max_stages(64) > 8 ? 8 : tc_max_stages(64))));
  { const char *v = getenv("DUALAR_TC_FUSE_NORM"); e->tc->fuse_norm = v ? atoi(v) : 0; }      // 2: statistics from the producing GEMM's epilogue (see ColBufs::ssq)
  e->tc->ws_bytes = (size_t)48 << 20;
  if ((rc = dev_alloc(e, e->tc->ws, e->tc->ws_bytes / 4)) || (rc = dev_alloc(e, e->tc->tickets, 8192)) || (rc = dev_alloc(e, e->tc->err, 4))) return rc;
  e->tc->ws_own = e->tc->ws; e->tc->tickets_own = e->tc->tickets;
  if ((rc = alloc_cols(e, e->tc->pf, 512, 1, false))) return rc;      // columns per prefill chunk: one chain of ~225 kernels per chunk, so prompts of up to 512 positions take one
  { const char *v = getenv("DUALAR_TC_CTAS"); if (v && atoi(v) > 0) e->tc->target_ctas = atoi(v); }
  { const char *v = getenv("DUALAR_TC_MINKB"); if (v && atoi(v) > 0) e->tc->min_kb = atoi(v); }
  { const char *v = getenv("DUALAR_ATTN_MMA"); e->tc->attn_mma = !(v && v[0] == '0'); }
  { const char *v = getenv("DUALAR_ATTN_CLUSTER"); e->tc->attn_cluster = !(v && v[0] == '0'); }
  { const char *v = getenv("DUALAR_TC_CLUSTER"); e->tc->cluster_reduce = !(v && v[0] == '0'); }
  { const char *v = getenv("DUALAR_ATTN_FUSE_POST"); e->tc->attn_fuse_post = !(v && v[0] == '0'); }
  { const char *v = getenv("DUALAR_ATTN_TPS"); if (v && atoi(v) > 0) e->tc->attn_tiles_per_split = atoi(v); }
  { const char *v = getenv("DUALAR_TC_KSPLIT"); if (v) e->tc->ksplit_override = atoi(v); }
  { const char *v = getenv("DUALAR_TC_STAGES"); if (v) e->tc->stages_override = atoi(v); }
  { const char *v = getenv("DUALAR_TC_PREFILL_BN"); if (v) e->tc->bn_override = atoi(v); }
  if (e->c.head_dim < 32 || e->c.head_dim > 128 || (e->c.head_dim & (e->c.head_dim - 1)))
    return fail(DUALAR_EINVAL, "the tensor-core path needs head_dim 32, 64 or 128 (a lane owns head_dim / 32 output dims and head_dim / 4 score dims in the attention kernel)");
  CU(cudaFuncSetAttribute(b_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b_attn_smem(e->c.n_head / e->c.n_local_heads, e->c.head_dim)));
  CU(cudaFuncSetAttribute(b_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b_attn_mma_smem(e->c.n_head / e->c.n_local_heads, e->c.head_dim)));
  CU(cudaFuncSetAttribute(b_fast_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)b_fast_attn_smem(e->c.fast_n_head, e->c.fast_n_local_heads, e->c.fast_head_dim, e->c.num_codebooks)));
  e->tc->ready = true;
  return 0;
}

// Y[n][r] = W[r][:] . X[n][:] on the tensor cores (gemm_tc.cuh); xcap = rows of the X buffer
// norm_w != nullptr: X is the UN-normalised activation and the kernel applies RMSNorm(norm_w) itself (BN = 32 only, see tc_can_fuse_norm)
static bool tc_can_fuse_norm(const dualar_engine *e, int BN, int K) { return e->tc->fuse_norm && BN == 32 && K % 256 == 0 && K <= 4096; }
// mode 2: the GEMM that writes `out` leaves the RMSNorm statistics of its output; the consumer's operand staging reads them
static float *ssq_out_for(dualar_engine *e, ColBufs &c, int BN, int rows, int epi, const bf16 *out) {
  if (e->tc->fuse_norm != 2 || BN != 32 || epi == TE_SWIGLU || rows % DA_TC_BM || rows / DA_TC_BM > DA_SSQ_LD) return nullptr;
  auto it = c.ssq.find(out); if (it == c.ssq.end()) return nullptr;
  c.ssq_valid.insert(out); return it->second;
}
static const float *ssq_in_for(dualar_engine *e, ColBufs &c, int BN, int K, const bf16 *x) {
  if (e->tc->fuse_norm != 2 || !tc_can_fuse_norm(e, BN, K) || !c.ssq_valid.count(x)) return nullptr;
  return c.ssq.at(x);
}
static int tc_gemm(dualar_engine *e, const bf16 *W, int rows, int K, const bf16 *X, int xcap, int ncols, int BN, int epi, const bf16 *bias,
                   const bf16 *res, bf16 *out, int w_keep, cudaStream_t s, int &count, const bf16 *norm_w = nullptr, bool prefill = false,
                   const float *ssq_in = nullptr, float *ssq_out = nullptr) {
  const CUtensorMap *mw, *mx; int rc;
  if (prefill) {
    // prefill: parallelism comes from COLUMN tiles, never from split-K -- a K split of a 256-column tile writes and re-reads
    // 128 x 256 fp32 partials per CTA, more bytes than the weights themselves; and with the K range whole, a position's result does
    // not depend on how the prompt was cut into chunks (prefix reuse prefills a short tail and must reproduce the full prefill).
    // Take the widest tile that still yields ~100 CTAs.
    const int rt0 = (rows + DA_TC_BM - 1) / DA_TC_BM;
    const int ks0 = (rt0 < 16 && K / DA_TC_BK >= 16) ? 4 : 1;
    for (BN = 256; BN > 32; BN >>= 1) if (ncols > BN / 2 && rt0 * ks0 * ((ncols + BN - 1) / BN) >= 96) break;
    if (e->tc->bn_override > 0) BN = e->tc->bn_override;
  }
  if ((rc = tc_map(e, W, rows, K, DA_TC_BM, &mw)) || (rc = tc_map(e, X, xcap, K, BN, &mx))) return rc;
  const int rt = (rows + DA_TC_BM - 1) / DA_TC_BM, ct = (ncols + BN - 1) / BN, tiles = rt * ct, nkb = K / DA_TC_BK;
  int ks = 1;
  if (!prefill && tiles < 64) {
    // with the cluster reduction a split is cheap: down to 2 k-blocks per CTA.  Measured (ms per step, 32 slots / 4 x 32 slots):
    // 128 CTAs of >= 4 k-blocks 3.44 / 5.74, 128 of >= 2: 3.38 / 5.77, 192 of >= 2: 3.26 / 6.05.  ONE rule for every mode: the K
    // partition fixes the summation order, and a request's bits must not depend on how many groups run beside it
    const int want = e->tc->target_ctas ? e->tc->target_ctas : 128;
    const int min_kb = e->tc->min_kb ? e->tc->min_kb : (e->tc->cluster_reduce ? 2 : 4);
    ks = want / tiles; if (ks > nkb / min_kb) ks = nkb / min_kb; if (ks > 8) ks = 8; if (ks < 1) ks = 1;
    if (e->tc->cluster_reduce) while (ks & (ks - 1)) --ks;      // the cluster reduction deals 128 / ks rows to every split
  }
  // prefill: a K split only for the matrices with very few row tiles (wo, w2: 8), chosen from the MATRIX alone so that a position's
  // summation order does not depend on how the prompt was chunked
  if (prefill && rt < 16 && nkb >= 16) ks = 4;
  if (!prefill && e->tc->ksplit_override > 0 && tiles < 64) { ks = e->tc->ksplit_override; if (ks > nkb) ks = nkb; }
  while (ks > 1 && (size_t)tiles * ks * BN * DA_TC_BM * 4 > e->tc->ws_bytes) --ks;
  if (tiles > 8192) return fail(DUALAR_EINVAL, "too many GEMM tiles (%d)", tiles);
  GemmTcArgs a; memset(&a, 0, sizeof(a));
  a.rows = rows; a.K = K; a.ncols = ncols; a.epi = epi; a.ld_out = epi == TE_SWIGLU ? rows / 2 : rows; a.w_keep = w_keep;
  a.bias = bias; a.res = res; a.out = out; a.ws = e->tc->ws; a.tickets = e->tc->tickets; a.err = e->tc->err;
  a.ssq_out = ssq_out; a.ssq_ld = rows / DA_TC_BM; a.ssq_in = ssq_in; a.ssq_n = K / DA_TC_BM;
  // many tiles: two CTAs per SM hide each other's set-up; few tiles: deep ring per CTA -- unless several request groups run
  // concurrently: then shared memory is better spent on co-resident CTAs of the other groups (4 x 32 slots: 7.02 ms per round with 4
  // stages against 7.26 with 8; 8 x 32: 10.9 against 12.7)
  int st = (tiles > 2 * e->sms || e->tc->concurrent_groups) ? 4 : 8;
  if (e->tc->stages_override > 0) st = e->tc->stages_override;
  const int mxs = tc_max_stages(BN); if (st > mxs) st = mxs;
  const int nk_per = (nkb + ks - 1) / ks; if (st > nk_per) st = nk_per < 2 ? 2 : nk_per;
  a.stages = st;
  const dim3 grid(rt, ct, ks), block(DA_TC_THREADS);
  if (BRec *r = e->tc->rec) {
    if (norm_w || prefill) return fail(DUALAR_EINVAL, "the persistent step records plain decode GEMMs only");
    BPhase &p = r->add(BP_GEMM, grid);
    p.map_w = r->map_index(mw); p.map_x = r->map_index(mx); p.u.gemm = a;
    return 0;
  }
  if (norm_w) {
    if (!tc_can_fuse_norm(e, BN, K) || ct != 1) return fail(DUALAR_EINVAL, "fused norm needs BN 32 and one column tile");
    a.xraw = X; a.norm_w = norm_w; a.eps = e->c.norm_eps;
    size_t smx = gemm_tc_smem_xn(BN, st, nk_per);
    while (smx > 200 * 1024 && st > 2) { --st; smx = gemm_tc_smem_xn(BN, st, nk_per); }
    if (smx > 200 * 1024) return fail(DUALAR_EINVAL, "fused-norm operand does not fit shared memory (K %d, splits %d)", K, ks);
    a.stages = st;
    CU(launch_k(gemm_tc_kernel<32, true>, grid, block, smx, s, *mw, *mx, a)); ++count;
    return 0;
  }
  // K splits of a tile as a thread-block cluster reducing through distributed shared memory (tiles of <= 64 columns: the receive buffer
  // sits behind the ring; power-of-two splits; same bits as the workspace path, so prefill chunking stays irrelevant; no
  // statistics emission, whose butterfly assumes the 16 row groups of a tile in one CTA)
  if (e->tc->cluster_reduce && (ks == 2 || ks == 4 || ks == 8) && (BN == 32 || BN == 64) && !ssq_out) {
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = gemm_tc_smem_cl(BN, st); cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = (unsigned)ks;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_use_pdl ? 2 : 1;
    if (BN == 32) CU(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<32, false, true>, *mw, *mx, a));
    else CU(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<64, false, true>, *mw, *mx, a));
    ++count;
    return 0;
  }
  const size_t smem = gemm_tc_smem(BN, st);
  switch (BN) {
    case 32: CU(launch_k(gemm_tc_kernel<32>, grid, block, smem, s, *mw, *mx, a)); break;
    case 64: CU(launch_k(gemm_tc_kernel<64>, grid, block, smem, s, *mw, *mx, a)); break;
    case 128: CU(launch_k(gemm_tc_kernel<128>, grid, block, smem, s, *mw, *mx, a)); break;
    case 256: CU(launch_k(gemm_tc_kernel<256>, grid, block, smem, s, *mw, *mx, a)); break;
    default: return fail(DUALAR_EINVAL, "BN %d", BN);
  }
  ++count;
  return 0;
}

static int bn_for(int ncols) { return ncols <= 32 ? 32 : ncols <= 64 ? 64 : ncols <= 128 ? 128 : 256; }

struct KvTarget { bf16 *const *kc, *const *vc; long long slot_stride; int S; };   // per-layer cache bases, elements between slots, cache length

// embedding + the slow stack over `ncols` columns; leaves the un-normalised last-layer output in c.x
static int enqueue_slow_cols(dualar_engine *e, ColBufs &c, int ncols, int BN, const KvTarget &kv, const PosSrc &pos, const TokSrc &tok,
                             cudaStream_t s, int &count, bool prefill = false) {
  const dualar_config &cf = e->c;
  const int qkv_rows = (cf.n_head + 2 * cf.n_local_heads) * cf.head_dim, qd = cf.n_head * cf.head_dim;
  int rc;
  { BEmbedArgs a; memset(&a, 0, sizeof(a));
    a.emb = e->emb; a.cb_emb = e->cb_emb; a.x = c.x; a.dim = cf.dim; a.vocab = cf.vocab_size; a.codebook_size = cf.codebook_size; a.num_codebooks = cf.num_codebooks;
    a.sem_begin = cf.semantic_begin_id; a.sem_end = cf.semantic_end_id; a.scale_cb = cf.scale_codebook_embeddings; a.cpu_sem = e->cpu_sem; a.ncols = ncols;
    a.inv_sqrt = (float)(1.0 / sqrt((double)(cf.num_codebooks + 1))); a.sqrt_c = (float)sqrt((double)(cf.num_codebooks + 1)); a.tok = tok; a.err = e->tc->err;
    if (BRec *r = e->tc->rec) r->add(BP_EMBED, dim3(ncols)).u.embed = a;
    else { CU(launch_k(b_embed_kernel, dim3(ncols), dim3(128), 0, s, a)); ++count; } }
  auto norm = [&](const bf16 *x, const bf16 *w, bf16 *out, int K) -> int {
    BNormArgs a{x, w, out, K, ncols, cf.norm_eps};
    if (BRec *r = e->tc->rec) { r->add(BP_NORM, dim3((ncols + 7) / 8)).u.norm = a; return 0; }      // one warp per column, 8 compute warps
    CU(launch_k(b_rmsnorm_kernel, dim3((ncols + 3) / 4), dim3(128), 0, s, a)); ++count; return 0; };
  c.ssq_valid.clear();      // the embedding kernel leaves no statistics; nothing carries over from the previous step
  // y = W . RMSNorm(x): a separate norm kernel in front of a TMA-fed GEMM (default), or the GEMM normalises its own operand
  // (DUALAR_TC_FUSE_NORM=1: statistics recomputed by every CTA; =2: statistics left by the GEMM that produced x, where there is one)
  auto normed_gemm = [&](const bf16 *W, int rows, int K, const bf16 *x, const bf16 *nw, bf16 *xn, int epi, const bf16 *bias, bf16 *out, int w_keep) -> int {
    const int mode = (prefill || e->tc->rec) ? 0 : e->tc->fuse_norm;
    if (mode == 1 && tc_can_fuse_norm(e, BN, K)) return tc_gemm(e, W, rows, K, x, c.cap, ncols, BN, epi, bias, nullptr, out, w_keep, s, count, nw);
    if (mode == 2) { if (const float *sq = ssq_in_for(e, c, BN, K, x)) return tc_gemm(e, W, rows, K, x, c.cap, ncols, BN, epi, bias, nullptr, out, w_keep, s, count, nw, false, sq); }
    int r2 = norm(x, nw, xn, K); if (r2) return r2;
    return tc_gemm(e, W, rows, K, xn, c.cap, ncols, BN, epi, bias, nullptr, out, w_keep, s, count, nullptr, prefill);
  };
  for (int l = 0; l < cf.n_layer; ++l) {
    LayerW &L = e->slow[l];
    if ((rc = normed_gemm(L.wqkv, qkv_rows, cf.dim, c.x, L.attn_norm, c.xn, TE_STORE, L.bqkv, c.qkv, 0))) return rc;
    // decode: q/k-norm + RoPE + the KV row write happen inside the attention kernel (one cache per column); prefill: the columns are
    // positions of ONE request that read each other's rows, so those must be in the cache before the attention kernel starts
    const bool fuse_post = !prefill && e->tc->attn_fuse_post;
    if (!fuse_post) { BQkvPostArgs a; memset(&a, 0, sizeof(a));
      a.qkv = c.qkv; a.kc = kv.kc[l]; a.vc = kv.vc[l]; a.slot_stride = kv.slot_stride; a.rope = e->rope; a.qn = L.qn; a.kn = L.kn;
      a.nh = cf.n_head; a.nkv = cf.n_local_heads; a.hd = cf.head_dim; a.S = kv.S; a.ncols = ncols; a.eps = cf.norm_eps; a.pos = pos;
      const dim3 g(ncols, (cf.n_head + 2 * cf.n_local_heads + 7) / 8);
      if (BRec *r = e->tc->rec) r->add(BP_QKV_POST, g).u.post = a;
      else { CU(launch_k(b_qkv_post_kernel, g, dim3(256), (size_t)8 * cf.head_dim * sizeof(float), s, a)); ++count; } }
    { BAttnArgs a; memset(&a, 0, sizeof(a));
      a.qkv = c.qkv; a.kc = kv.kc[l]; a.vc = kv.vc[l]; a.slot_stride = kv.slot_stride; a.nh = cf.n_head; a.nkv = cf.n_local_heads; a.hd = cf.head_dim; a.S = kv.S;
      a.ncols = ncols; a.nsplit_max = c.nsplit; a.sf = (float)sqrt(1.0 / sqrt((double)cf.head_dim)); a.part_o = c.part_o; a.part_ml = c.part_ml;
      a.tickets = c.attn_tickets; a.y = c.y; a.pos = pos; a.err = e->tc->err;
      a.fuse_post = fuse_post; a.rope = e->rope; a.qn = L.qn; a.kn = L.kn; a.eps = cf.norm_eps;
      const dim3 g(c.nsplit, cf.n_local_heads, ncols);
      // KV splits as a cluster merging through distributed shared memory: pays when the GPU is nearly empty (8 slots: 2.81 against 3.01 ms
      // per step, a split per tile) and costs when it is not (32 slots: 3.68 against 3.40; 4 x 32: 7.1 against 5.8 -- four co-scheduled CTAs
      // per (request, kv head), most of them idle).  So: engines with at most 8 slots only; everything larger splits from 8 tiles on and
      // merges through the ticket, whatever the grouping (the partition fixes the merge order, hence the bits)
      const bool small = e->tc->attn_cluster && e->batch_total > 0 && e->batch_total <= 8 && c.nsplit > 1 && c.nsplit <= DA_B_MAXSPLIT && !prefill;
      a.tiles_per_split = e->tc->attn_tiles_per_split ? e->tc->attn_tiles_per_split : (small ? 1 : 8);
      a.cluster_merge = small && !e->tc->rec;
      const int Gq = cf.n_head / cf.n_local_heads;
      // tensor-core attention: decode only (one tensor map per layer over the group's whole cache; prefill targets one slot's cache)
      a.use_mma = !prefill && kv.slot_stride > 0 && e->tc->attn_mma && Gq <= 8 && (cf.head_dim == 64 || cf.head_dim == 128);
      const size_t asmem = a.use_mma ? b_attn_mma_smem(Gq, cf.head_dim) : b_attn_smem(Gq, cf.head_dim);
      const CUtensorMap *mkp = nullptr, *mvp = nullptr;
      if (a.use_mma) {
        const long long rows_total = (long long)c.cap * (kv.slot_stride / cf.head_dim);
        if ((rc = tc_map(e, kv.kc[l], (int)rows_total, cf.head_dim, DA_TILE, &mkp)) || (rc = tc_map(e, kv.vc[l], (int)rows_total, cf.head_dim, DA_TILE, &mvp))) return rc;
      }
      auto launch_attn = [&](cudaLaunchConfig_t &cfg) -> cudaError_t {
        return a.use_mma ? cudaLaunchKernelEx(&cfg, b_attn_mma_kernel, *mkp, *mvp, a) : cudaLaunchKernelEx(&cfg, b_attn_kernel, a);
      };
      if (BRec *r = e->tc->rec) { BPhase &ph = r->add(BP_ATTN, g); ph.u.attn = a; if (a.use_mma) { ph.map_w = r->map_index(mkp); ph.map_x = r->map_index(mvp); } }
      else if (a.cluster_merge) {
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = g; cfg.blockDim = dim3(DA_ATTN_THREADS); cfg.dynamicSmemBytes = asmem; cfg.stream = s;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)c.nsplit; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = g_use_pdl ? 2 : 1;
        CU(launch_attn(cfg)); ++count;
      }
      else {
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = g; cfg.blockDim = dim3(DA_ATTN_THREADS); cfg.dynamicSmemBytes = asmem; cfg.stream = s;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = g_use_pdl ? 1 : 0;
        CU(launch_attn(cfg)); ++count;
      } }
    const bool emit = !prefill && !e->tc->rec;
    if ((rc = tc_gemm(e, L.wo, cf.dim, qd, c.y, c.cap, ncols, BN, TE_RESIDUAL, L.bo, c.x, c.h, 0, s, count, nullptr, prefill, nullptr,
                      emit ? ssq_out_for(e, c, BN, cf.dim, TE_RESIDUAL, c.h) : nullptr))) return rc;
    if ((rc = normed_gemm(L.w13, 2 * cf.intermediate_size, cf.dim, c.h, L.ffn_norm, c.xn, TE_SWIGLU, nullptr, c.act, 0))) return rc;
    if ((rc = tc_gemm(e, L.w2, cf.dim, cf.intermediate_size, c.act, c.cap, ncols, BN, TE_RESIDUAL, nullptr, c.h, c.x, 0, s, count, nullptr, prefill, nullptr,
                      emit ? ssq_out_for(e, c, BN, cf.dim, TE_RESIDUAL, c.x) : nullptr))) return rc;
  }
  return 0;
}

// ---- prefill: KV rows of prompt positions [t0, t1) of the request whose token rows live in `seq` (device, row stride seq_stride) ----
static int tc_prefill(dualar_engine *e, const KvTarget &kv, const int *seq, int seq_stride, int t0, int t1, cudaStream_t s) {
  int rc = tc_init(e); if (rc) return rc;
  ColBufs &c = e->tc->pf;
  int count = 0;
  for (int t = t0; t < t1; t += c.cap) {
    const int ncols = t1 - t < c.cap ? t1 - t : c.cap;
    PosSrc pos{nullptr, 0, t};
    TokSrc tok{seq + t, 1, seq_stride};
    if ((rc = enqueue_slow_cols(e, c, ncols, bn_for(ncols), kv, pos, tok, s, count, true))) return rc;
  }
  e->prefill_launches = count;
  return 0;
}

// ---- batched decode ------------------------------------------------------------------------------------------------------------------
// slow sampler (inference.py:103-113): repetition penalty + chunk maxima, then exact nucleus statistics and the draw in the last CTA per request
static int enqueue_slow_sampler(dualar_engine *e, cudaStream_t s2, int &count) {
  dualar_batch &b = *e->batch; ColBufs &c = b.c; const dualar_config &cf = e->c; const int B = b.B, R = cf.num_codebooks + 1;
  { BHeadArgs a{c.logits, e->batch_keep_raw ? c.logits_raw : nullptr, c.cmax, cf.vocab_size, b.nchunk, R, b.st};
    CU(launch_k(b_head_stats_kernel, dim3(b.nchunk, B), dim3(512), 0, s2, a)); ++count; }
  { BSelectArgs a; memset(&a, 0, sizeof(a));
    a.logits = c.logits; a.cmax = c.cmax; a.V = cf.vocab_size; a.nchunk = b.nchunk; a.delta = e->delta; a.cand = c.cand; a.fast_emb = e->fast_emb; a.fast_x = c.fin;
    a.fast_dim = cf.fast_dim; a.codebook_size = cf.codebook_size; a.sem_begin = cf.semantic_begin_id; a.st = b.st;
    CU(launch_k(b_select_kernel, dim3(b.nchunk, B), dim3(512), (size_t)(192 * 8 + 34 * 8 + 80 * 4 + 64), s2, a)); ++count; }
  return 0;
}

static int enqueue_batch_step(dualar_engine *e, cudaStream_t s, int &count) {
  dualar_batch &b = *e->batch; ColBufs &c = b.c;
  const dualar_config &cf = e->c;
  const int B = b.B, BN = b.BN, R = cf.num_codebooks + 1;
  const long long sst = (long long)(sizeof(DAState) / sizeof(int));
  int rc;
  KvTarget kv{b.kc.data(), b.vc.data(), b.slot_stride, b.Sb};
  PosSrc pos{&b.st->pos, sst, 0};
  TokSrc tok{b.st->tok_in, sst, 1};
  if ((rc = enqueue_slow_cols(e, c, B, BN, kv, pos, tok, s, count))) return rc;
  auto norm = [&](const bf16 *x, const bf16 *w, bf16 *out, int K) -> int {
    BNormArgs a{x, w, out, K, B, cf.norm_eps};
    if (BRec *r = e->tc->rec) { r->add(BP_NORM, dim3((B + 7) / 8)).u.norm = a; return 0; }
    CU(launch_k(b_rmsnorm_kernel, dim3((B + 3) / 4), dim3(128), 0, s, a)); ++count; return 0; };
  // LM head + slow sampler (llama.py:446-451, inference.py:103-113) on a SECOND stream: they need only the slow hidden state, and so
  // does pass 0 of the fast stack (whose logits the reference discards, inference.py:121-122) -- the two branches are independent
  // until pass 1 consumes the sampled semantic id.  Fork / join with events (captured as graph edges).
  BRec *rec = e->tc->rec;
  const bool fork = e->batch_fork && !rec;
  cudaStream_t s2 = fork ? b.side_stream : s;
  if (fork) { CU(cudaEventRecord(b.ev_fork, s)); CU(cudaStreamWaitEvent(s2, b.ev_fork, 0)); }
  { BNormArgs a{c.x, e->norm, c.xn, cf.dim, B, cf.norm_eps};
    if (rec) rec->add(BP_NORM, dim3((B + 7) / 8)).u.norm = a;
    else { CU(launch_k(b_rmsnorm_kernel, dim3((B + 3) / 4), dim3(128), 0, s2, a)); ++count; } }
  if ((rc = tc_gemm(e, cf.tie_word_embeddings ? e->emb : e->out_w, cf.vocab_size, cf.dim, c.xn, c.cap, B, BN, TE_STORE, nullptr, nullptr, c.logits, 0, s2, count))) return rc;
  if (rec) rec->split = (int)rec->ph.size();      // the slow sampler's two kernels run between the two persistent launches
  else if ((rc = enqueue_slow_sampler(e, s2, count))) return rc;
  if (fork) CU(cudaEventRecord(b.ev_join, s2));
  // fast AR: pass 0 consumes the slow hidden state, pass k >= 1 the embedding of codebook k-1 (inference.py:121-149)
  const int fqkv_rows = (cf.fast_n_head + 2 * cf.fast_n_local_heads) * cf.fast_head_dim, fqd = cf.fast_n_head * cf.fast_head_dim;
  const bool emit = !rec;
  c.ssq_valid.erase(c.fin);      // written by the samplers (an embedding row): no statistics
  auto normed_gemm = [&](const bf16 *W, int rows, int K, const bf16 *x, const bf16 *nw, bf16 *xn, int epi, const bf16 *bias, bf16 *out) -> int {
    const int mode = rec ? 0 : e->tc->fuse_norm;
    if (mode == 1 && tc_can_fuse_norm(e, BN, K)) return tc_gemm(e, W, rows, K, x, c.cap, B, BN, epi, bias, nullptr, out, 1, s, count, nw);
    if (mode == 2) { if (const float *sq = ssq_in_for(e, c, BN, K, x)) return tc_gemm(e, W, rows, K, x, c.cap, B, BN, epi, bias, nullptr, out, 1, s, count, nw, false, sq); }
    int r2 = norm(x, nw, xn, K); if (r2) return r2;
    return tc_gemm(e, W, rows, K, xn, c.cap, B, BN, epi, bias, nullptr, out, 1, s, count);
  };
  if (e->fpi_w) {      // hidden_states = fast_project_in(x)   (llama.py:590)
    if ((rc = tc_gemm(e, e->fpi_w, cf.fast_dim, cf.dim, c.x, c.cap, B, BN, TE_STORE, e->fpi_b, nullptr, c.fpi, 1, s, count, nullptr, false, nullptr,
                      emit ? ssq_out_for(e, c, BN, cf.fast_dim, TE_STORE, c.fpi) : nullptr))) return rc;
  }
  for (int p = 0; p < cf.num_codebooks; ++p) {
    const bf16 *in = p == 0 ? (e->fpi_w ? c.fpi : c.x) : c.fin;
    for (int l = 0; l < cf.n_fast_layer; ++l) {
      LayerW &L = e->fast[l];
      bf16 *out = c.fx[l & 1];
      if ((rc = normed_gemm(L.wqkv, fqkv_rows, cf.fast_dim, in, L.attn_norm, c.fxn, TE_STORE, L.bqkv, c.fqkv))) return rc;
      { BFastAttnArgs a; memset(&a, 0, sizeof(a));
        a.qkv = c.fqkv; a.kc = b.fkc[l]; a.vc = b.fvc[l]; a.slot_stride = b.fslot_stride; a.rope = e->fast_rope; a.qn = L.qn; a.kn = L.kn;
        a.nh = cf.fast_n_head; a.nkv = cf.fast_n_local_heads; a.hd = cf.fast_head_dim; a.ncb = cf.num_codebooks; a.p = p; a.ncols = B;
        a.eps = cf.norm_eps; a.scale = (float)(1.0 / sqrt((double)cf.fast_head_dim)); a.y = c.fy;
        if (rec) rec->add(BP_FAST_ATTN, dim3(B)).u.fattn = a;
        else { CU(launch_k(b_fast_attn_kernel, dim3(B), dim3(256), b_fast_attn_smem(a.nh, a.nkv, a.hd, a.ncb), s, a)); ++count; } }
      if ((rc = tc_gemm(e, L.wo, cf.fast_dim, fqd, c.fy, c.cap, B, BN, TE_RESIDUAL, L.bo, in, c.fh, 1, s, count, nullptr, false, nullptr,
                        emit ? ssq_out_for(e, c, BN, cf.fast_dim, TE_RESIDUAL, c.fh) : nullptr))) return rc;
      if ((rc = normed_gemm(L.w13, 2 * cf.fast_intermediate_size, cf.fast_dim, c.fh, L.ffn_norm, c.fxn, TE_SWIGLU, nullptr, c.fact))) return rc;
      if ((rc = tc_gemm(e, L.w2, cf.fast_dim, cf.fast_intermediate_size, c.fact, c.cap, B, BN, TE_RESIDUAL, nullptr, c.fh, out, 1, s, count, nullptr, false, nullptr,
                        emit ? ssq_out_for(e, c, BN, cf.fast_dim, TE_RESIDUAL, out) : nullptr))) return rc;
      in = out;
    }
    if (p == 0) {              // logits of pass 0 are discarded by the reference (inference.py:122)
      if (fork) CU(cudaStreamWaitEvent(s, b.ev_join, 0));      // join: pass 1 starts from the embedding of the sampled id
      continue;
    }
    if ((rc = normed_gemm(e->fast_out, e->fv, cf.fast_dim, in, e->fast_norm, c.fxn, TE_STORE, nullptr, c.flogits))) return rc;
    { BFastSampleArgs a; memset(&a, 0, sizeof(a));
      a.logits = c.flogits; a.logits_raw = e->batch_keep_raw ? c.flogits_raw : nullptr; a.fv = e->fv; a.head = p; a.ncb = cf.num_codebooks; a.last_head = (p == cf.num_codebooks - 1);
      a.noise_off = (long long)cf.vocab_size + (long long)(p - 1) * e->fv; a.fast_emb = e->fast_emb; a.fast_x = c.fin; a.fast_dim = cf.fast_dim; a.codebook_size = cf.codebook_size;
      a.seq = b.seq; a.seq_slot_stride = (long long)R * b.Sb; a.seq_stride = b.Sb; a.im_end_id = cf.im_end_id; a.n_rows_tok = R; a.st = b.st;
      if (rec) rec->add(BP_FAST_SAMPLE, dim3(B)).u.fsample = a;
      else { CU(launch_k(b_fast_sample_kernel, dim3(B), dim3(256), b_fast_sample_smem(), s, a)); ++count; } }
  }
  return 0;
}

// ---- the persistent step: record the sequence once, upload the phase table and tensor maps, capture two cooperative launches ----------
template <int BN> static cudaError_t launch_bstep(dualar_engine *e, const BStepArgs &k, cudaStream_t s) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(e->sms); cfg.blockDim = dim3(DA_BS_THREADS); cfg.dynamicSmemBytes = e->batch->bs_smem; cfg.stream = s;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, bstep_kernel<BN>, k);
}
static int enqueue_bstep(dualar_engine *e, int ph0, int ph1, cudaStream_t s, int &count) {
  dualar_batch &b = *e->batch;
  CU(cudaMemsetAsync(b.gbar, 0, sizeof(unsigned int), s));
  BStepArgs k; k.phases = b.d_phases + ph0; k.n_phases = ph1 - ph0; k.maps = b.d_maps; k.gbar = b.gbar; k.err = e->tc->err; k.stages = b.bs_stages;
  { const char *v = getenv("DUALAR_BS_NO_PREFETCH"); k.no_prefetch = v && v[0] == '1'; }
  k.tl = b.d_tl ? b.d_tl + 2 * ph0 : nullptr;
  switch (b.BN) {
    case 32: CU(launch_bstep<32>(e, k, s)); break;
    case 64: CU(launch_bstep<64>(e, k, s)); break;
    case 128: CU(launch_bstep<128>(e, k, s)); break;
    default: return fail(DUALAR_EINVAL, "BN %d", b.BN);
  }
  ++count;
  return 0;
}
static int enqueue_batch_step_persistent(dualar_engine *e, cudaStream_t s, int &count) {
  dualar_batch &b = *e->batch; int rc;
  if ((rc = enqueue_bstep(e, 0, b.split, s, count))) return rc;             // embedding, slow stack, final norm, LM head
  if ((rc = enqueue_slow_sampler(e, s, count))) return rc;
  return enqueue_bstep(e, b.split, b.n_phases, s, count);                   // the fast passes and their samplers
}
static int build_persistent_step(dualar_engine *e) {
  dualar_batch &b = *e->batch; const dualar_config &cf = e->c;
  if (e->tc->fuse_norm) return 0;      // the fused-norm experiment exists on the per-kernel path only
  int coop = 0; CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device));
  if (!coop) return 0;
  BRec rec; e->tc->rec = &rec;
  int n = 0; int rc = enqueue_batch_step(e, e->cap_stream, n);
  e->tc->rec = nullptr;
  if (rc < 0) return rc;
  if (rec.split < 0) return fail(DUALAR_ESTATE, "recorded step has no sampler split");
  size_t body = std::max(b_attn_smem(cf.n_head / cf.n_local_heads, cf.head_dim), b_attn_mma_smem(cf.n_head / cf.n_local_heads, cf.head_dim));
  body = std::max(body, b_fast_attn_smem(cf.fast_n_head, cf.fast_n_local_heads, cf.fast_head_dim, cf.num_codebooks));
  body = std::max(body, b_fast_sample_smem());
  body = std::max(body, (size_t)8 * cf.head_dim * sizeof(float));
  body = (body + 1023) & ~(size_t)1023;
  const size_t stg = (size_t)b.BN * DA_TC_BM * 4; if (body < stg) body = stg;
  int st = (int)((227 * 1024 - 2048 - body) / (DA_TC_A_BYTES + (size_t)b.BN * 128));
  if (st > DA_TC_MAX_STAGES) st = DA_TC_MAX_STAGES;
  { const char *v = getenv("DUALAR_BS_STAGES"); if (v && atoi(v) >= 2 && atoi(v) < st) st = atoi(v); }
  if (st < 2) return 0;      // does not fit: the per-kernel graph remains the only path
  b.bs_stages = st; b.bs_smem = bstep_smem(b.BN, st, body);
  switch (b.BN) {
    case 32: CU(cudaFuncSetAttribute(bstep_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.bs_smem)); break;
    case 64: CU(cudaFuncSetAttribute(bstep_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.bs_smem)); break;
    case 128: CU(cudaFuncSetAttribute(bstep_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.bs_smem)); break;
    default: return 0;      // 256 columns: per-kernel path only
  }
  b.n_phases = (int)rec.ph.size(); b.split = rec.split;
  if ((rc = dev_alloc(e, b.d_phases, rec.ph.size())) || (rc = dev_alloc(e, b.d_maps, rec.maps.size())) || (rc = dev_alloc(e, b.gbar, 4))) return rc;
  CU(cudaMemcpy(b.d_phases, rec.ph.data(), rec.ph.size() * sizeof(BPhase), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(b.d_maps, rec.maps.data(), rec.maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  for (const BPhase &p : rec.ph) b.kinds.push_back(p.kind | ((p.gx * p.gy * p.gz) << 8));
  { const char *v = getenv("DUALAR_BS_TIMELINE"); if (v && v[0] == '1') { if ((rc = dev_alloc(e, b.d_tl, 2 * rec.ph.size()))) return rc; } }
  // dry run, then capture
  n = 0;
  if ((rc = enqueue_batch_step_persistent(e, e->cap_stream, n)) < 0) return rc;
  CU(cudaStreamSynchronize(e->cap_stream));
  { int gerr[4] = {0, 0, 0, 0}; CU(cudaMemcpy(gerr, e->tc->err, sizeof(gerr), cudaMemcpyDeviceToHost));
    if (gerr[0]) return fail(DUALAR_EDEVICE, "persistent batched step: device fault flag %d on the dry run (5-9: a bounded wait timed out; phase %d / producer phase %d of %d, split %d)",
                             gerr[0], gerr[1], gerr[2], b.n_phases, b.split); }
  cudaGraph_t g;
  CU(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
  n = 0;
  rc = enqueue_batch_step_persistent(e, e->cap_stream, n);
  cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &g);
  if (rc < 0) return rc;
  if (ce != cudaSuccess) return fail(DUALAR_ECUDA, "persistent batch graph capture failed: %s", cudaGetErrorString(ce));
  CU(cudaGraphInstantiate(&b.g_step_p, g, 0));
  CU(cudaGraphDestroy(g));
  b.launches_p = n;
  return 0;
}
static int batch_select_path(dualar_engine *e) {
  if (e->groups.empty()) return 0;
  int want = e->batch_persistent;
  if (want < 0) { const char *v = getenv("DUALAR_BATCH_PERSIST"); want = v ? (v[0] != '0') : 0; }
  for (dualar_batch *b : e->groups) if (want && !b->g_step_p) return fail(DUALAR_ESTATE, "the persistent batched step is not available for this configuration");
  for (dualar_batch *b : e->groups) b->persistent = want != 0;
  return 0;
}

// one request group: its own activation buffers, KV caches, sampling state, split-K workspace and step graph(s)
static int batch_group_init(dualar_engine *e, int n_slots, int slot_seq_len, bool own_stream) {
  const dualar_config &cf = e->c; int rc;
  dualar_batch *b = new dualar_batch(); e->groups.push_back(b); e->batch = b;
  b->B = n_slots; b->BN = bn_for(n_slots); b->Sb = slot_seq_len;
  const int R = cf.num_codebooks + 1;
  // split-KV: up to 4 splits per (request, kv head), whatever the group size -- the partition fixes the order of the softmax merge, and
  // a request's bits must not depend on the grouping.  With the cluster merge a split costs about a microsecond, so even two tiles are split
  int nsplit = 4;
  { const char *v = getenv("DUALAR_BATCH_NSPLIT"); if (v) nsplit = atoi(v); }
  if ((rc = alloc_cols(e, b->c, b->BN, nsplit, true))) return rc;
  b->slot_stride = (long long)cf.n_local_heads * b->Sb * cf.head_dim;
  b->fslot_stride = (long long)cf.fast_n_local_heads * cf.num_codebooks * cf.fast_head_dim;
  b->kc.resize(cf.n_layer); b->vc.resize(cf.n_layer); b->fkc.resize(cf.n_fast_layer); b->fvc.resize(cf.n_fast_layer);
  for (int l = 0; l < cf.n_layer; ++l) if ((rc = dev_alloc(e, b->kc[l], (size_t)b->slot_stride * b->BN)) || (rc = dev_alloc(e, b->vc[l], (size_t)b->slot_stride * b->BN))) return rc;
  for (int l = 0; l < cf.n_fast_layer; ++l) if ((rc = dev_alloc(e, b->fkc[l], (size_t)b->fslot_stride * b->BN)) || (rc = dev_alloc(e, b->fvc[l], (size_t)b->fslot_stride * b->BN))) return rc;
  if ((rc = dev_alloc(e, b->st, (size_t)b->BN)) || (rc = dev_alloc(e, b->seq, (size_t)b->BN * R * b->Sb))) return rc;
  if (!own_stream) { b->ws = e->tc->ws_own; b->tickets = e->tc->tickets_own; }      // a single group: its prefills run on the same stream as its steps
  else if ((rc = dev_alloc(e, b->ws, e->tc->ws_bytes / 4)) || (rc = dev_alloc(e, b->tickets, 8192))) return rc;
  CU(cudaMallocHost((void **)&b->h_st, sizeof(DAState)));
  CU(cudaMallocHost((void **)&b->h_seq, (size_t)R * b->Sb * sizeof(int)));
  CU(cudaMallocHost((void **)&b->h_act, sizeof(DAState) * (size_t)n_slots));
  CU(cudaMallocHost((void **)&b->h_prompt, (size_t)R * b->Sb * sizeof(int)));
  CU(cudaEventCreateWithFlags(&b->ev_prompt, cudaEventDisableTiming));
  b->prompt_len.assign(n_slots, 0); b->max_gen.assign(n_slots, 0); b->open.assign(n_slots, 0); b->known_done.assign(n_slots, 0);
  CU(cudaStreamCreateWithFlags(&b->side_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
  if (own_stream) {
    CU(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&b->ev_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->ev_pf, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&b->ev_act, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->ev_rel, cudaEventDisableTiming));
  }
  if ((rc = dev_alloc(e, b->st_idle, 1))) return rc;
  { DAState idle; memset(&idle, 0, sizeof(idle)); idle.pos = -1; CU(cudaMemcpy(b->st_idle, &idle, sizeof(idle), cudaMemcpyHostToDevice)); }
  // slow sampler: (chunks x requests) CTAs of 512 threads scan the logits; the register-heavy sampler code runs one CTA per SM, so
  // more CTAs than SMs only adds waves
  b->nchunk = e->sms / n_slots; if (b->nchunk < 1) b->nchunk = 1; if (b->nchunk > 32) b->nchunk = 32;
  // the group's GEMMs are enqueued with the group's split-K workspace
  e->tc->ws = b->ws; e->tc->tickets = b->tickets; e->tc->concurrent_groups = own_stream;
  // dry run (configures attributes, surfaces launch errors), then capture
  int n = 0;
  if ((rc = enqueue_batch_step(e, e->cap_stream, n)) < 0) return rc;
  CU(cudaStreamSynchronize(e->cap_stream));
  { int gerr = 0; CU(cudaMemcpy(&gerr, e->tc->err, sizeof(int), cudaMemcpyDeviceToHost));
    if (gerr) return fail(DUALAR_EDEVICE, "batched step: device fault flag %d on the dry run", gerr); }
  cudaGraph_t g;
  CU(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
  n = 0;
  rc = enqueue_batch_step(e, e->cap_stream, n);
  cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &g);
  if (rc < 0) return rc;
  if (ce != cudaSuccess) return fail(DUALAR_ECUDA, "batch graph capture failed: %s", cudaGetErrorString(ce));
  CU(cudaGraphInstantiate(&b->g_step, g, 0));
  CU(cudaGraphDestroy(g));
  b->launches = n;
  if ((rc = build_persistent_step(e))) return rc;
  e->tc->ws = e->tc->ws_own; e->tc->tickets = e->tc->tickets_own; e->tc->concurrent_groups = false;
  // the dry run advanced nothing (every slot is idle: loop_mode 0) but wrote KV row 0 and sampler scratch; start clean
  for (int i = 0; i < b->BN; ++i) CU(cudaMemcpy(b->st + i, b->st_idle, sizeof(DAState), cudaMemcpyDeviceToDevice));
  for (int l = 0; l < cf.n_layer; ++l) { CU(cudaMemset(b->kc[l], 0, (size_t)b->slot_stride * b->BN * 2)); CU(cudaMemset(b->vc[l], 0, (size_t)b->slot_stride * b->BN * 2)); }
  CU(cudaMemset(e->tc->err, 0, 4));
  CU(cudaDeviceSynchronize());
  return 0;
}

// Request slots live in GROUPS of `batch_group_slots` (option, default 32): a group is one batched step (one graph, its own buffers and
// stream); the groups of a step run CONCURRENTLY.  One step of one group is a chain of ~540 dependent kernels that leaves the GPU mostly
// idle (8 us per kernel, a few MB each), so several independent chains overlap almost for free: measured on B200 (s1-mini), 1 x 32 slots
// 7.3 k tok/s, 2 x 32 11.4 k, 4 x 32 17.1 k.  The weights are shared; slot s belongs to group s / group_slots.
extern "C" int dualar_batch_init(dualar_engine *e, int max_batch, int slot_seq_len) {
  if (!e) return fail(DUALAR_EINVAL, "null engine");
  if (!e->finalized) return fail(DUALAR_ESTATE, "dualar_finalize has not been called");
  if (e->batch) return fail(DUALAR_ESTATE, "batch slots already exist");
  const dualar_config &cf = e->c;
  if (max_batch < 1 || max_batch > 1024) return fail(DUALAR_EINVAL, "max_batch must be in [1, 1024]");
  if (slot_seq_len <= 0 || slot_seq_len > cf.max_seq_len) slot_seq_len = cf.max_seq_len;
  slot_seq_len = (slot_seq_len + 7) / 8 * 8;
  if (cf.n_head / cf.n_local_heads > DA_MAX_G) return fail(DUALAR_EINVAL, "GQA group too large");
  CU(cudaSetDevice(e->device));
  int rc = tc_init(e); if (rc) return rc;
  { const char *v = getenv("DUALAR_BATCH_FORK"); e->batch_fork = !(v && v[0] == '0'); }
  int gs = e->batch_group_slots;
  { const char *v = getenv("DUALAR_BATCH_GROUP_SLOTS"); if (v && atoi(v) > 0) gs = atoi(v); }
  if (gs < 1) gs = 1; if (gs > 128) gs = 128;
  const int n_groups = (max_batch + gs - 1) / gs;
  e->group_slots = gs; e->batch_total = max_batch;
  for (int g = 0; g < n_groups; ++g) {
    const int n = std::min(gs, max_batch - g * gs);
    if ((rc = batch_group_init(e, n, slot_seq_len, n_groups > 1))) return rc;
  }
  if (n_groups > 1) { CU(cudaEventCreateWithFlags(&e->ev_groups_go, cudaEventDisableTiming)); CU(cudaStreamCreateWithFlags(&e->pf_stream, cudaStreamNonBlocking)); }
  e->batch = e->groups[0];
  return batch_select_path(e);
}

// global slot -> its group and the slot index inside it
static dualar_batch *group_of(dualar_engine *e, int slot, int *local) {
  if (slot < 0 || slot >= e->batch_total) return nullptr;
  *local = slot % e->group_slots;
  return e->groups[slot / e->group_slots];
}

extern "C" int dualar_batch_prefill(dualar_engine *e, int slot, const int32_t *prompt, int T, int max_new, float temperature, float top_p, float rep,
                                    uint64_t seed, const void *noise, void *stream) {
  if (!e || !prompt) return fail(DUALAR_EINVAL, "null argument");
  if (!e->batch) return fail(DUALAR_ESTATE, "dualar_batch_init has not been called");
  dualar_batch *bp = group_of(e, slot, &slot);
  if (!bp) return fail(DUALAR_EINVAL, "slot out of range");
  dualar_batch &b = *bp; const dualar_config &cf = e->c; const int R = cf.num_codebooks + 1;
  if (T < 1) return fail(DUALAR_EINVAL, "empty prompt");
  if (T >= b.Sb) return fail(DUALAR_EINVAL, "Input sequence length %d exceeds the slot length %d", T, b.Sb);
  if (max_new <= 0 || T + max_new > b.Sb) max_new = b.Sb - T;
  CU(cudaSetDevice(e->device));
  // Several groups: the prefill runs ASYNCHRONOUSLY on the engine's prefill stream while the other groups (and this group's other slots)
  // keep decoding; the slot is switched on by a copy of its state on the GROUP's stream, behind an event -- i.e. between two steps of the
  // group, never in the middle of one.  One group: everything on the caller's stream, synchronously.
  const bool async = b.stream != nullptr;
  cudaStream_t s = async ? e->pf_stream : (cudaStream_t)stream;
  // the prompt staging buffer is free once the previous prompt's copy has run (it runs on the prefill stream, not behind decode steps); a
  // finished request parks its slot (position -1), so the steps still in flight for the group do not touch the cache being filled
  if (async) {
    if (b.prompt_pending) { CU(cudaEventSynchronize(b.ev_prompt)); b.prompt_pending = false; }
    if (b.rel_pending) { CU(cudaStreamWaitEvent(s, b.ev_rel, 0)); b.rel_pending = false; }
  }
  else CU(cudaStreamSynchronize(s));
  int *seq = b.seq + (size_t)slot * R * b.Sb;
  for (int r = 0; r < R; ++r) memcpy(b.h_prompt + (size_t)r * T, prompt + (size_t)r * T, (size_t)T * sizeof(int));
  CU(cudaMemcpy2DAsync(seq, (size_t)b.Sb * sizeof(int), b.h_prompt, (size_t)T * sizeof(int), (size_t)T * sizeof(int), R, cudaMemcpyHostToDevice, s));
  if (async) { CU(cudaEventRecord(b.ev_prompt, s)); b.prompt_pending = true; }
  std::vector<bf16 *> kc(cf.n_layer), vc(cf.n_layer);
  for (int l = 0; l < cf.n_layer; ++l) { kc[l] = b.kc[l] + (size_t)slot * b.slot_stride; vc[l] = b.vc[l] + (size_t)slot * b.slot_stride; }
  KvTarget kv{kc.data(), vc.data(), 0, b.Sb};
  int rc = tc_prefill(e, kv, seq, b.Sb, 0, T - 1, s); if (rc) return rc;
  // the slot joins the batch at the LAST prompt position: the next batched step produces its first token (no penalty, inference.py:353-362)
  DAState *h = b.h_act + slot; memset(h, 0, sizeof(*h));      // per slot: the copy below may sit behind a burst of decode steps
  h->pos = T - 1; h->max_gen = max_new; h->prompt_len = T; h->loop_mode = 1; h->temperature = temperature; h->top_p = top_p; h->rep_penalty = rep;
  h->seed = seed; h->cpu_sem = e->cpu_sem; h->noise = (const bf16 *)noise; h->noise_stride = (long long)cf.vocab_size + (long long)(cf.num_codebooks - 1) * e->fv;
  for (int r = 0; r < R; ++r) h->tok_in[r] = prompt[(size_t)r * T + (T - 1)];
  if (async) {
    CU(cudaEventRecord(b.ev_pf, s));
    CU(cudaStreamWaitEvent(b.stream, b.ev_pf, 0));
    CU(cudaMemcpyAsync(b.st + slot, h, sizeof(*h), cudaMemcpyHostToDevice, b.stream));
    CU(cudaEventRecord(b.ev_act, b.stream));
    b.act_pending = true;
  } else {
    CU(cudaMemcpyAsync(b.st + slot, h, sizeof(*h), cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));      // the pinned staging buffers are reused by the next call
  }
  b.prompt_len[slot] = T; b.max_gen[slot] = max_new; b.open[slot] = 1; b.known_done[slot] = 0;
  return 0;
}

extern "C" int dualar_batch_decode(dualar_engine *e, int n_steps, void *stream) {
  if (!e) return fail(DUALAR_EINVAL, "null engine");
  if (!e->batch) return fail(DUALAR_ESTATE, "dualar_batch_init has not been called");
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (e->groups.size() == 1) {
    dualar_batch &b = *e->groups[0];
    cudaGraphExec_t g = b.persistent ? b.g_step_p : b.g_step;
    for (int i = 0; i < n_steps; ++i) CU(cudaGraphLaunch(g, s));
    return 0;
  }
  // several groups: fork from the caller's stream, one chain of graph launches per group (enqueued round-robin so that no group's queue
  // runs dry while the host is busy with another's), join back.  A group with no open slot only burns idle steps: skip it.
  CU(cudaEventRecord(e->ev_groups_go, s));
  std::vector<dualar_batch *> live;
  for (dualar_batch *b : e->groups) { bool any = false; for (char o : b->open) any = any || o; if (any) live.push_back(b); }
  for (dualar_batch *b : live) CU(cudaStreamWaitEvent(b->stream, e->ev_groups_go, 0));
  for (int i = 0; i < n_steps; ++i)
    for (dualar_batch *b : live) CU(cudaGraphLaunch(b->persistent ? b->g_step_p : b->g_step, b->stream));
  // option batch_decode_join = 0 (the pipelined serving loop): no join here -- dualar_batch_read and dualar_batch_collect make the caller's
  // stream wait for the groups they look at, so the host can collect finished requests and enqueue prefills while the next burst runs
  for (dualar_batch *b : live) {
    CU(cudaEventRecord(b->ev_done, b->stream)); b->decode_pending = true;
    if (e->batch_decode_join) CU(cudaStreamWaitEvent(s, b->ev_done, 0));
  }
  return 0;
}

extern "C" int dualar_batch_collect(dualar_engine *e, int slot, int32_t *out, int cap, int *n_tokens, int *finished, void *stream) {
  if (!e || !n_tokens) return fail(DUALAR_EINVAL, "null argument");
  if (!e->batch) return fail(DUALAR_ESTATE, "dualar_batch_init has not been called");
  const int gslot = slot;
  dualar_batch *bp = group_of(e, slot, &slot);
  if (!bp || !bp->open[slot]) return fail(DUALAR_ESTATE, "slot %d holds no request", gslot);
  dualar_batch &b = *bp; const int R = e->c.num_codebooks + 1;
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  // a request the host has SEEN finished (dualar_batch_read "done") is stable: no need to wait for the steps enqueued since
  if (!b.known_done[slot]) {
    if (b.act_pending) CU(cudaStreamWaitEvent(s, b.ev_act, 0));
    if (b.decode_pending) CU(cudaStreamWaitEvent(s, b.ev_done, 0));
  }
  CU(cudaMemcpyAsync(b.h_st, b.st + slot, sizeof(DAState), cudaMemcpyDeviceToHost, s));
  int gerr = 0; CU(cudaMemcpyAsync(&gerr, e->tc->err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (b.h_st->err || gerr) return fail(DUALAR_EDEVICE, "device fault flag: slot %d, kernels %d (1: token id out of range, 2: bulk-copy wait timed out, 3: code >= codebook_size, 5-7: tensor-core pipeline wait timed out)", b.h_st->err, gerr);
  const int n = b.h_st->n_gen;
  *n_tokens = n;
  if (finished) *finished = b.h_st->done;
  if (out && n > 0) {
    if (cap < n) return fail(DUALAR_EINVAL, "output capacity %d < %d generated columns", cap, n);
    const int *seq = b.seq + (size_t)slot * R * b.Sb;
    CU(cudaMemcpy2DAsync(b.h_seq, (size_t)n * sizeof(int), seq + b.prompt_len[slot], (size_t)b.Sb * sizeof(int), (size_t)n * sizeof(int), R, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (int r = 0; r < R; ++r) memcpy(out + (size_t)r * cap, b.h_seq + (size_t)r * n, (size_t)n * sizeof(int));
  }
  return 0;
}

extern "C" int dualar_batch_release(dualar_engine *e, int slot) {
  if (!e || !e->batch) return fail(DUALAR_ESTATE, "dualar_batch_init has not been called");
  dualar_batch *bp = group_of(e, slot, &slot);
  if (!bp) return fail(DUALAR_EINVAL, "slot out of range");
  dualar_batch &b = *bp;
  CU(cudaSetDevice(e->device));
  if (b.stream) {
    // several groups: no device-wide synchronisation (other groups' prefills may be in flight); the idle state is written on the
    // group's stream, i.e. after every step already enqueued for the group and before any later one
    CU(cudaMemcpyAsync(b.st + slot, b.st_idle, sizeof(DAState), cudaMemcpyDeviceToDevice, b.stream));
    // an UNFINISHED request released early may still be writing its KV rows in steps already enqueued: a later prefill of the slot waits
    // for them; a finished one is parked (position -1) and needs no such wait
    if (!b.known_done[slot]) { CU(cudaEventRecord(b.ev_rel, b.stream)); b.rel_pending = true; }
  } else {
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(b.st + slot, b.st_idle, sizeof(DAState), cudaMemcpyDeviceToDevice));      // idle: loop_mode 0, position -1
  }
  b.open[slot] = 0;
  return 0;
}

// one group's share of a named buffer: copies min(nbytes, what the group holds) and reports what it holds; *per_slot = 0 for scalars
static int batch_read_group(dualar_engine *e, dualar_batch &b, const char *name, void *dst, int64_t nbytes, cudaStream_t s, int64_t *held, int *per_slot) {
  const dualar_config &cf = e->c;
  const void *src = nullptr; int64_t avail = 0; size_t pitch = 0, width = 0;
  *per_slot = 1;
  if (!strcmp(name, "slow_logits")) { src = b.c.logits; avail = (int64_t)b.B * cf.vocab_size * 2; }
  else if (!strcmp(name, "slow_logits_raw")) { src = b.c.logits_raw; avail = (int64_t)b.B * cf.vocab_size * 2; }
  else if (!strcmp(name, "hidden")) { src = b.c.x; avail = (int64_t)b.B * cf.dim * 2; }
  else if (!strcmp(name, "fast_logits")) { src = b.c.flogits_raw; avail = (int64_t)b.B * (cf.num_codebooks - 1) * e->fv * 2; }
  else if (!strcmp(name, "state")) { src = b.st; avail = (int64_t)b.B * sizeof(DAState); }
  else if (!strcmp(name, "tokens")) { src = b.st->tok_out; pitch = sizeof(DAState); width = (size_t)(cf.num_codebooks + 1) * 4; avail = (int64_t)b.B * width; }
  else if (!strcmp(name, "positions")) { src = &b.st->pos; pitch = sizeof(DAState); width = 4; avail = (int64_t)b.B * width; }
  else if (!strcmp(name, "done")) { src = &b.st->done; pitch = sizeof(DAState); width = 4; avail = (int64_t)b.B * width; }
  else if (!strcmp(name, "n_gen")) { src = &b.st->n_gen; pitch = sizeof(DAState); width = 4; avail = (int64_t)b.B * width; }
  else {
    *per_slot = 0;
    if (!strcmp(name, "launches")) { int n = 0; for (dualar_batch *g : e->groups) n += g->persistent ? g->launches_p : g->launches; *(int *)dst = n; return nbytes >= 4 ? 0 : fail(DUALAR_EINVAL, "4 bytes needed"); }
    if (!strcmp(name, "groups")) { *(int *)dst = (int)e->groups.size(); return nbytes >= 4 ? 0 : fail(DUALAR_EINVAL, "4 bytes needed"); }
    if (!strcmp(name, "bstep_kinds")) { if (nbytes > (int64_t)b.kinds.size() * 4) return fail(DUALAR_EINVAL, "too many bytes"); memcpy(dst, b.kinds.data(), (size_t)nbytes); return 0; }
    if (!strcmp(name, "bstep_phases")) { *(int *)dst = b.n_phases; return nbytes >= 4 ? 0 : fail(DUALAR_EINVAL, "4 bytes needed"); }
    if (!strcmp(name, "bstep_timeline")) {
      if (!b.d_tl) return fail(DUALAR_ESTATE, "DUALAR_BS_TIMELINE=1 was not set at batch_init");
      if (nbytes > (int64_t)b.n_phases * 16) return fail(DUALAR_EINVAL, "too many bytes");
      CU(cudaMemcpyAsync(dst, b.d_tl, (size_t)nbytes, cudaMemcpyDeviceToHost, s)); CU(cudaStreamSynchronize(s)); return 0;
    }
    return fail(DUALAR_EINVAL, "unknown batch buffer '%s'", name);
  }
  *held = avail;
  const int64_t n = nbytes < avail ? nbytes : avail;
  if (pitch) CU(cudaMemcpy2DAsync(dst, width, src, pitch, width, (size_t)(n / (int64_t)width), cudaMemcpyDeviceToHost, s));
  else CU(cudaMemcpyAsync(dst, src, (size_t)n, cudaMemcpyDeviceToHost, s));
  return 0;
}

// per-slot buffers are returned for all slots in slot order (group after group)
extern "C" int dualar_batch_read(dualar_engine *e, const char *name, void *dst, int64_t nbytes, void *stream) {
  if (!e || !name || !dst) return fail(DUALAR_EINVAL, "null argument");
  if (!e->batch) return fail(DUALAR_ESTATE, "dualar_batch_init has not been called");
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  for (dualar_batch *b : e->groups) {
    if (b->act_pending) CU(cudaStreamWaitEvent(s, b->ev_act, 0));
    if (b->decode_pending) CU(cudaStreamWaitEvent(s, b->ev_done, 0));
  }
  int64_t off = 0;
  for (dualar_batch *b : e->groups) {
    int64_t held = 0; int per_slot = 0;
    int rc = batch_read_group(e, *b, name, (char *)dst + off, nbytes - off, s, &held, &per_slot);
    if (rc || !per_slot) return rc;
    off += held;
    if (off >= nbytes) break;
  }
  if (off < nbytes) return fail(DUALAR_EINVAL, "batch buffer '%s' holds %lld bytes, %lld requested", name, (long long)off, (long long)nbytes);
  CU(cudaStreamSynchronize(s));
  for (dualar_batch *b : e->groups) { b->decode_pending = false; b->act_pending = false; }      // everything enqueued so far has run
  if (!strcmp(name, "done")) {      // remember which requests the host has seen finished (dualar_batch_collect / release rely on it)
    const int32_t *d = (const int32_t *)dst; int64_t i = 0, n = nbytes / 4;
    for (dualar_batch *b : e->groups) for (int sl = 0; sl < b->B && i < n; ++sl, ++i) if (d[i]) b->known_done[sl] = 1;
  }
  return 0;
}

static void batch_destroy(dualar_engine *e) {
  for (dualar_batch *b : e->groups) {
    if (b->g_step) cudaGraphExecDestroy(b->g_step);
    if (b->g_step_p) cudaGraphExecDestroy(b->g_step_p);
    if (b->side_stream) cudaStreamDestroy(b->side_stream);
    if (b->stream) cudaStreamDestroy(b->stream);
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    if (b->ev_done) cudaEventDestroy(b->ev_done);
    if (b->ev_pf) cudaEventDestroy(b->ev_pf);
    if (b->ev_act) cudaEventDestroy(b->ev_act);
    if (b->ev_rel) cudaEventDestroy(b->ev_rel);
    if (b->h_st) cudaFreeHost(b->h_st);
    if (b->h_seq) cudaFreeHost(b->h_seq);
    if (b->h_act) cudaFreeHost(b->h_act);
    if (b->h_prompt) cudaFreeHost(b->h_prompt);
    if (b->ev_prompt) cudaEventDestroy(b->ev_prompt);
    delete b;
  }
  e->groups.clear(); e->batch = nullptr;
  if (e->ev_groups_go) { cudaEventDestroy(e->ev_groups_go); e->ev_groups_go = nullptr; }
  if (e->pf_stream) { cudaStreamDestroy(e->pf_stream); e->pf_stream = nullptr; }
  delete e->tc; e->tc = nullptr;
}
