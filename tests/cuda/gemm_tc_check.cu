// gemm_tc_check.cu -- stand-alone check + timing of the tcgen05 / TMEM / TMA GEMM (fish_tts_b200/csrc/gemm_tc.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/gemm_tc_check tests/cuda/gemm_tc_check.cu
// Part 1: every (BN, epilogue, split-K) combination against a one-thread-per-output fp32 reference.  With small-integer data
// every partial sum is exact in fp32, so the comparison is BIT-EXACT whatever the summation order: any descriptor / swizzle /
// TMEM-lane mistake shows up as a mismatch.  With Gaussian data the two fp32 orders may differ by one bf16 ulp.
// Part 2: timing of the decode-step shapes (s1-mini, 32 columns) launched back to back with programmatic dependent launch.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../../fish_tts_b200/csrc/gemm_tc.cuh"

using namespace da;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static uint16_t f2b(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float b2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

__global__ void ref_kernel(const bf16 *W, const bf16 *X, const bf16 *bias, const bf16 *res, bf16 *out, int rows, int K, int ncols, int epi, int ld) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
  if (r >= rows || n >= ncols) return;
  auto dot = [&](int rr) { float acc = 0.f; for (int k = 0; k < K; ++k) acc = fmaf(bf2f(W[(size_t)rr * K + k]), bf2f(X[(size_t)n * K + k]), acc); return acc; };
  if (epi == TE_SWIGLU) {
    if (r & 1) return;
    const float g = rbf(dot(r) + (bias ? bf2f(bias[r]) : 0.f)), u = rbf(dot(r + 1) + (bias ? bf2f(bias[r + 1]) : 0.f));
    const float sg = rbf(g / (1.0f + expf(-g)));
    out[(size_t)n * ld + (r >> 1)] = f2bf(__fmul_rn(sg, u));
  } else {
    float y = rbf(dot(r) + (bias ? bf2f(bias[r]) : 0.f));
    if (epi == TE_RESIDUAL) y = bf2f(res[(size_t)n * ld + r]) + y;
    out[(size_t)n * ld + r] = f2bf(y);
  }
}

static int stages_for(int BN, int want) {
  int mx = (int)((227 * 1024 - 2048 - 1024) / (DA_TC_A_BYTES + BN * 128));
  if (mx > DA_TC_MAX_STAGES) mx = DA_TC_MAX_STAGES;
  return want < mx ? want : mx;
}

template <int BN> static cudaError_t launch(const CUtensorMap &mw, const CUtensorMap &mx, const GemmTcArgs &a, dim3 grid, cudaStream_t s, bool pdl) {
  static bool configured = false;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_tc_smem(BN, stages_for(BN, DA_TC_MAX_STAGES)));
    if (ce != cudaSuccess) return ce;
    configured = true;
  }
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(DA_TC_THREADS); cfg.dynamicSmemBytes = gemm_tc_smem(BN, a.stages); cfg.stream = s;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN>, mw, mx, a);
}
static cudaError_t launch_bn(int BN, const CUtensorMap &mw, const CUtensorMap &mx, const GemmTcArgs &a, dim3 grid, cudaStream_t s, bool pdl) {
  switch (BN) {
    case 32: return launch<32>(mw, mx, a, grid, s, pdl);
    case 64: return launch<64>(mw, mx, a, grid, s, pdl);
    case 128: return launch<128>(mw, mx, a, grid, s, pdl);
    default: return launch<256>(mw, mx, a, grid, s, pdl);
  }
}
struct Case { int rows, K, ncols, BN, ksplit, epi, bias, ints; };

int main(int argc, char **argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  int *d_err; CK(cudaMalloc(&d_err, 4)); CK(cudaMemset(d_err, 0, 4));
  long long *d_dbg; CK(cudaMalloc(&d_dbg, (size_t)2048 * 8 * 8)); CK(cudaMemset(d_dbg, 0, (size_t)2048 * 8 * 8));
  auto stamps = [&](const char *what, int nctas) {      // per-CTA clock64 stamps of the last launch: median durations of the kernel's sections
    std::vector<long long> h((size_t)nctas * 8); CK(cudaMemcpy(h.data(), d_dbg, h.size() * 8, cudaMemcpyDeviceToHost));
    auto med = [&](int a, int b) { std::vector<long long> v; for (int c = 0; c < nctas; ++c) if (h[c * 8 + a] && h[c * 8 + b]) v.push_back(h[c * 8 + b] - h[c * 8 + a]); if (v.empty()) return -1.0; std::sort(v.begin(), v.end()); return (double)v[v.size() / 2]; };
    printf("  stamps %s (median cycles over %d CTAs): set-up %.0f | dep-wait since start %.0f | first stage landed since start %.0f | last MMA issued since first stage %.0f | accumulator since first stage %.0f | epilogue %.0f | total %.0f\n",
           what, nctas, med(0, 1), med(0, 2), med(0, 3), med(3, 4), med(3, 5), med(5, 6), med(0, 6));
  };
  float *d_ws; CK(cudaMalloc(&d_ws, (size_t)64 << 20));
  unsigned *d_tk; CK(cudaMalloc(&d_tk, 1 << 20)); CK(cudaMemset(d_tk, 0, 1 << 20));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  int fails = 0;

  std::vector<Case> cases = {
    {256, 256, 5, 32, 1, TE_STORE, 0, 1},        {256, 256, 5, 32, 1, TE_STORE, 0, 0},
    {1024, 1024, 32, 32, 1, TE_STORE, 1, 1},     {1024, 2048, 32, 32, 4, TE_RESIDUAL, 1, 1},
    {6144, 1024, 32, 32, 1, TE_SWIGLU, 0, 1},    {1024, 3072, 17, 32, 8, TE_RESIDUAL, 0, 1},
    {4096, 1024, 223, 256, 1, TE_STORE, 0, 1},   {4096, 1024, 223, 256, 1, TE_STORE, 1, 0},
    {1024, 2048, 223, 256, 2, TE_RESIDUAL, 0, 0},{6144, 1024, 100, 128, 1, TE_SWIGLU, 0, 0},
    {1024, 3072, 60, 64, 3, TE_RESIDUAL, 1, 0},  {640, 256, 9, 32, 1, TE_STORE, 0, 0},      // 640 rows: partial last tile
    {2048, 1024, 300, 256, 1, TE_STORE, 0, 1},   {2048, 1024, 300, 128, 2, TE_SWIGLU, 0, 0},  // two / three column tiles
    {102048, 1024, 8, 32, 1, TE_STORE, 0, 0},                                                 // the 1.5-shape head: 797.25 tiles
  };
  for (const Case &c : cases) {
    const int nalloc = ((c.ncols + c.BN - 1) / c.BN) * c.BN;
    const int ld = c.epi == TE_SWIGLU ? c.rows / 2 : c.rows;
    std::vector<uint16_t> hW((size_t)c.rows * c.K), hX((size_t)nalloc * c.K, 0), hB(c.rows), hR((size_t)c.ncols * ld);
    uint32_t rng = 12345u + c.rows * 7 + c.K * 3 + c.ncols;
    auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return (rng >> 8) * (1.0f / 16777216.0f); };
    auto val = [&](float scale) { if (c.ints) return (float)((int)(rnd() * 9.f) - 4) * (scale >= 1.f ? 1.f : 0.25f); float g = 0.f; for (int i = 0; i < 4; ++i) g += rnd() - 0.5f; return g * scale; };
    for (auto &v : hW) v = f2b(val(c.ints ? 1.f : 0.06f));
    for (int n = 0; n < c.ncols; ++n) for (int k = 0; k < c.K; ++k) hX[(size_t)n * c.K + k] = f2b(val(1.f));
    for (auto &v : hB) v = f2b(val(c.ints ? 1.f : 0.5f));
    for (auto &v : hR) v = f2b(val(1.f));
    bf16 *dW, *dX, *dB, *dR, *dO, *dRef;
    CK(cudaMalloc(&dW, hW.size() * 2)); CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dR, hR.size() * 2));
    CK(cudaMalloc(&dO, (size_t)c.ncols * ld * 2)); CK(cudaMalloc(&dRef, (size_t)c.ncols * ld * 2));
    CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dR, hR.data(), hR.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0xFF, (size_t)c.ncols * ld * 2)); CK(cudaMemset(dRef, 0, (size_t)c.ncols * ld * 2));
    CUtensorMap mw, mx;
    if (!tc_make_map(&mw, dW, c.rows, c.K, DA_TC_BM) || !tc_make_map(&mx, dX, nalloc, c.K, c.BN)) { printf("tensor map encode failed\n"); return 2; }
    GemmTcArgs a; memset(&a, 0, sizeof(a));
    a.rows = c.rows; a.K = c.K; a.ncols = c.ncols; a.stages = stages_for(c.BN, 6); a.epi = c.epi; a.ld_out = ld; a.w_keep = 0;
    a.bias = c.bias ? dB : nullptr; a.res = dR; a.out = dO; a.ws = d_ws; a.tickets = d_tk; a.err = d_err;
    dim3 grid((c.rows + DA_TC_BM - 1) / DA_TC_BM, nalloc / c.BN, c.ksplit);
    CK(launch_bn(c.BN, mw, mx, a, grid, s, false));
    ref_kernel<<<dim3((c.rows + 127) / 128, c.ncols), 128, 0, s>>>(dW, dX, a.bias, dR, dRef, c.rows, c.K, c.ncols, c.epi, ld);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    int herr = 0; CK(cudaMemcpy(&herr, d_err, 4, cudaMemcpyDeviceToHost));
    std::vector<uint16_t> o((size_t)c.ncols * ld), r((size_t)c.ncols * ld);
    CK(cudaMemcpy(o.data(), dO, o.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r.data(), dRef, r.size() * 2, cudaMemcpyDeviceToHost));
    size_t bad = 0, off1 = 0; double worst = 0.0; size_t first_bad = (size_t)-1;
    for (size_t i = 0; i < o.size(); ++i) {
      if (o[i] == r[i]) continue;
      const float fo = b2f(o[i]), fr = b2f(r[i]);
      const int d = abs((int)(o[i] & 0x7FFF) - (int)(r[i] & 0x7FFF));
      // one bf16 ulp of the LINEAR output (|y| up to ~4 here: 2^-6), which a residual add can leave on a much smaller result
      if ((o[i] >> 15) == (r[i] >> 15) && d <= 1) ++off1;
      else if (fabsf(fo - fr) <= 0.0157f + 4e-3f * fabsf(fr)) ++off1;
      else { ++bad; if (first_bad == (size_t)-1) first_bad = i; }
      if (fabs((double)fo - fr) > worst) worst = fabs((double)fo - fr);
    }
    const bool pass = herr == 0 && bad == 0 && (!c.ints || off1 == 0);
    printf("%s rows %6d K %4d ncols %3d BN %3d ksplit %d epi %d bias %d %s: exact %.4f%%, 1-ulp %zu, wrong %zu, worst |d| %.4g, err flag %d",
           pass ? "PASS" : "FAIL", c.rows, c.K, c.ncols, c.BN, c.ksplit, c.epi, c.bias, c.ints ? "ints " : "gauss",
           100.0 * (double)(o.size() - bad - off1) / (double)o.size(), off1, bad, worst, herr);
    if (bad) printf("  first wrong at n=%zu r=%zu: got %g want %g", first_bad / ld, first_bad % ld, b2f(o[first_bad]), b2f(r[first_bad]));
    printf("\n");
    if (!pass) ++fails;
    if (herr) { printf("device fault flag raised: stopping (a lost transaction means a descriptor / barrier bug)\n"); return 3; }
    cudaFree(dW); cudaFree(dX); cudaFree(dB); cudaFree(dR); cudaFree(dO); cudaFree(dRef);
  }

  // ---- timing: the GEMMs of one s1-mini slow layer + the LM head at 32 columns, back to back with PDL -----------------------------
  {
    struct Shape { const char *name; int rows, K, epi; };
    const Shape shapes[] = {{"wqkv", 4096, 1024, TE_STORE}, {"wo", 1024, 2048, TE_RESIDUAL}, {"w13", 6144, 1024, TE_SWIGLU}, {"w2", 1024, 3072, TE_RESIDUAL}, {"head", 155776, 1024, TE_STORE}};
    const int BN = 32, NL = 28;
    for (int variant = 0; variant < 5; ++variant) {
      // K splits per shape {wqkv, wo, w13, w2}: more CTAs stream in parallel (one SM pulls ~50 GB/s through TMA), the last split to
      // arrive reduces the partials
      static const int KS[5][4] = {{1, 4, 1, 4}, {2, 4, 2, 4}, {4, 4, 4, 4}, {4, 8, 4, 8}, {4, 8, 4, 12}};
      const int st_want = 11;
      // separate weights per layer so that nothing is L2-resident (28 x 42 MB), one shared activation buffer per K
      std::vector<bf16 *> W[4]; bf16 *Whead;
      for (int i = 0; i < 4; ++i) for (int l = 0; l < NL; ++l) { bf16 *p; CK(cudaMalloc(&p, (size_t)shapes[i].rows * shapes[i].K * 2)); CK(cudaMemset(p, 0x11, (size_t)shapes[i].rows * shapes[i].K * 2)); W[i].push_back(p); }
      CK(cudaMalloc(&Whead, (size_t)155776 * 1024 * 2)); CK(cudaMemset(Whead, 0x11, (size_t)155776 * 1024 * 2));
      bf16 *X[4], *O; for (int i = 0; i < 4; ++i) { CK(cudaMalloc(&X[i], (size_t)BN * 3072 * 2)); CK(cudaMemset(X[i], 0, (size_t)BN * 3072 * 2)); }
      CK(cudaMalloc(&O, (size_t)BN * 155776 * 2));
      std::vector<CUtensorMap> mw[4], mx(5); CUtensorMap mhead;
      for (int i = 0; i < 4; ++i) { mw[i].resize(NL); for (int l = 0; l < NL; ++l) tc_make_map(&mw[i][l], W[i][l], shapes[i].rows, shapes[i].K, DA_TC_BM); tc_make_map(&mx[i], X[i], BN, shapes[i].K, BN); }
      tc_make_map(&mhead, Whead, 155776, 1024, DA_TC_BM); tc_make_map(&mx[4], X[0], BN, 1024, BN);
      auto run = [&](bool with_head) {
        for (int l = 0; l < NL; ++l) for (int i = 0; i < 4; ++i) {
          GemmTcArgs a; memset(&a, 0, sizeof(a));
          a.rows = shapes[i].rows; a.K = shapes[i].K; a.ncols = BN; a.epi = shapes[i].epi; a.ld_out = a.epi == TE_SWIGLU ? a.rows / 2 : a.rows;
          const int ks = KS[variant][i];
          // ring no deeper than the CTA's own k-blocks: small CTAs leave room for the NEXT kernel's CTAs on the same SM, whose set-up and
          // weight prefetch then overlap this kernel (programmatic dependent launch)
          a.stages = stages_for(BN, std::min(st_want, (a.K / DA_TC_BK + ks - 1) / ks)); a.res = O; a.out = O; a.ws = d_ws; a.tickets = d_tk; a.err = d_err;
          CK(launch_bn(BN, mw[i][l], mx[i], a, dim3(a.rows / DA_TC_BM, 1, ks), s, true));
        }
        if (with_head) {
          GemmTcArgs a; memset(&a, 0, sizeof(a));
          a.rows = 155776; a.K = 1024; a.ncols = BN; a.epi = TE_STORE; a.ld_out = a.rows; a.stages = stages_for(BN, 4); a.out = O; a.ws = d_ws; a.tickets = d_tk; a.err = d_err;
          CK(launch_bn(BN, mhead, mx[4], a, dim3(155776 / DA_TC_BM, 1, 1), s, true));
        }
      };
      cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
      run(true); CK(cudaStreamSynchronize(s));
      { int herr0 = 0; CK(cudaMemcpy(&herr0, d_err, 4, cudaMemcpyDeviceToHost)); if (herr0) { printf("fault flag %d in the timing warm-up: stopping\n", herr0); return 3; } }
      if (fails) { printf("correctness failures above: skipping the timing\n"); break; }
      CK(cudaEventRecord(e0, s)); for (int it = 0; it < 5; ++it) run(false); CK(cudaEventRecord(e1, s));
      for (int it = 0; it < 5; ++it) run(true); CK(cudaEventRecord(e2, s));
      CK(cudaStreamSynchronize(s));
      float ms_layers, ms_all; CK(cudaEventElapsedTime(&ms_layers, e0, e1)); CK(cudaEventElapsedTime(&ms_all, e1, e2));
      ms_layers /= 5; ms_all /= 5;
      const double layer_bytes = 28.0 * (4096.0 * 1024 + 1024.0 * 2048 + 6144.0 * 1024 + 1024.0 * 3072) * 2, head_bytes = 155776.0 * 1024 * 2;
      int herr = 0; CK(cudaMemcpy(&herr, d_err, 4, cudaMemcpyDeviceToHost));
      printf("timing ksplit {wqkv %d, wo %d, w13 %d, w2 %d} stages<=%d: 112 layer GEMMs %.3f ms (%.2f us each, %.0f GB/s); head alone %.3f ms (%.0f GB/s); err %d\n",
             KS[variant][0], KS[variant][1], KS[variant][2], KS[variant][3], stages_for(BN, st_want), ms_layers, ms_layers * 1000 / 112, layer_bytes / ms_layers * 1e-6, ms_all - ms_layers, head_bytes / (ms_all - ms_layers) * 1e-6, herr);
      if (variant == 2 || variant == 0) for (int i = 0; i < 4; ++i) {
        GemmTcArgs a; memset(&a, 0, sizeof(a));
        a.rows = shapes[i].rows; a.K = shapes[i].K; a.ncols = BN; a.epi = shapes[i].epi; a.ld_out = a.epi == TE_SWIGLU ? a.rows / 2 : a.rows;
        const int ks = KS[variant][i];
        a.stages = stages_for(BN, st_want); a.res = O; a.out = O; a.ws = d_ws; a.tickets = d_tk; a.err = d_err; a.dbg = d_dbg;
        CK(cudaMemset(d_dbg, 0, (size_t)2048 * 8 * 8));
        CK(launch_bn(BN, mw[i][5], mx[i], a, dim3(a.rows / DA_TC_BM, 1, ks), s, false)); CK(cudaStreamSynchronize(s));
        stamps(shapes[i].name, a.rows / DA_TC_BM * ks);
      }
      for (int i = 0; i < 4; ++i) { for (auto p : W[i]) cudaFree(p); cudaFree(X[i]); }
      cudaFree(Whead); cudaFree(O);
    }
  }
  // ---- prefill-like: 223 columns, BN = 256 -----------------------------------------------------------------------------------------
  {
    const int BN = 256, T = 223;
    struct Shape { const char *name; int rows, K, epi, ks; };
    const Shape shapes[] = {{"wqkv", 4096, 1024, TE_STORE, 1}, {"wo", 1024, 2048, TE_RESIDUAL, 4}, {"w13", 6144, 1024, TE_SWIGLU, 1}, {"w2", 1024, 3072, TE_RESIDUAL, 4}};
    bf16 *W, *X, *O; CK(cudaMalloc(&W, (size_t)6144 * 3072 * 2)); CK(cudaMemset(W, 0x11, (size_t)6144 * 3072 * 2));
    CK(cudaMalloc(&X, (size_t)BN * 3072 * 2)); CK(cudaMemset(X, 0, (size_t)BN * 3072 * 2)); CK(cudaMalloc(&O, (size_t)BN * 6144 * 2));
    for (const Shape &sh : shapes) {
      CUtensorMap mw, mx; tc_make_map(&mw, W, sh.rows, sh.K, DA_TC_BM); tc_make_map(&mx, X, BN, sh.K, BN);
      GemmTcArgs a; memset(&a, 0, sizeof(a));
      a.rows = sh.rows; a.K = sh.K; a.ncols = T; a.epi = sh.epi; a.ld_out = a.epi == TE_SWIGLU ? a.rows / 2 : a.rows; a.stages = stages_for(BN, 4);
      a.res = O; a.out = O; a.ws = d_ws; a.tickets = d_tk; a.err = d_err;
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(launch_bn(BN, mw, mx, a, dim3(a.rows / DA_TC_BM, 1, sh.ks), s, true)); CK(cudaStreamSynchronize(s));
      CK(cudaEventRecord(e0, s)); for (int it = 0; it < 20; ++it) CK(launch_bn(BN, mw, mx, a, dim3(a.rows / DA_TC_BM, 1, sh.ks), s, true)); CK(cudaEventRecord(e1, s));
      CK(cudaStreamSynchronize(s));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 20;
      printf("prefill-like %s T=%d BN=%d ksplit %d: %.2f us (%.1f TFLOP/s)\n", sh.name, T, BN, sh.ks, ms * 1000, 2.0 * sh.rows * sh.K * T / ms * 1e-9);
      a.dbg = d_dbg; CK(cudaMemset(d_dbg, 0, (size_t)2048 * 8 * 8));
      CK(launch_bn(BN, mw, mx, a, dim3(a.rows / DA_TC_BM, 1, sh.ks), s, false)); CK(cudaStreamSynchronize(s));
      stamps(sh.name, a.rows / DA_TC_BM * sh.ks);
    }
  }
  printf(fails ? "RESULT: %d case(s) FAILED\n" : "RESULT: all cases passed\n", fails);
  return fails ? 1 : 0;
}
