"""Parity of the CUDA decode path (through the C-ABI) against the oracle -- runs on the B200 box.

Layers of evidence, from strict to statistical:
  1. the SAMPLER alone on identical logits is bit-exact against the reference sampler (index-ordered ties);
  2. teacher-forced steps (oracle trajectory + oracle KV state, "phase A" binding): per-step logits within the
     north-star tolerance; every sampled id equals what the reference sampler draws from OUR logits;
  3. the engine-owned loop (own prefill, own KV, device-side window / EOS) reproduces the oracle's `generate`;
  4. the fixtures of the UNMODIFIED reference (CPU) are met under `cpu_scalar_semantics`;
  5. at BASELINE sizes: size-independent properties (determinism, id ranges, EOS, idempotent replay).
bf16 note (DESIGN.md "numerics"): any two correct bf16 pipelines that differ only in fp32 summation order
(cuBLAS vs MKL vs ours) decorrelate to ~1-3 bf16 ulp after 28 layers; on s1-mini the reference's own CPU and
CUDA paths differ by 0.03-0.04 on semantic logits.  The full-size test therefore also measures that yardstick.
"""
import ctypes as C

import numpy as np
import pytest
import torch
from torch.nn.attention import SDPBackend, sdpa_kernel

from fish_tts_b200 import philox
from fish_tts_b200.config import s1_mini_config, tiny_config
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt
from helpers import bf16_ulp, logits_close, near_tie, variant_configs

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from fish_tts_b200 import capi
    from fish_tts_b200.engine import DualAREngine
    from gpu_common import TeacherForced, block_noise_source, build_pair, oracle_generate_tokenwise_prefill
from oracle import dualar_oracle as orc
from pathlib import Path

GOLD = Path(__file__).resolve().parent / "golden"
MODES = {"sampled": (0.7, 0.8, 1.1), "greedy": (0.7, 1e-9, 1.0), "hot": (1.0, 1.0, 1.5)}


def oracle_sample(cfg, logits_raw, head, window, T, p, rp, blk, device="cuda:0"):
    """the reference sampler (inference.py:30-80, index-ordered ties) on given raw logits of one head"""
    fv = min(1024, cfg.codebook_size)
    lg = logits_raw.to(device).clone().view(1, 1, -1)
    prev = None
    if window is not None:
        prev = window[:, 0] if head == 0 else window[head + 1]
        prev = prev.to(device)
    off = 0 if head == 0 else cfg.vocab_size + (head - 1) * fv
    noise = blk[off: off + lg.numel()]
    t = [torch.tensor(v, device=device, dtype=torch.float) for v in (T, p, rp)]
    tok, _ = orc.sample(lg, *t, prev, noise=orc.NoiseSource(lambda c, n: noise), stable_ties=True)
    return int(tok)


def check_step(cfg, o, T, p, rp, where, slow_tol=None, atol=2e-2, ulps=2.0):
    """one teacher-forced step: logits within tolerance up to the first token divergence; our ids = reference sampler on OUR logits"""
    logits_close(o["my_slow"], o["ref_slow"], cfg, f"{where} slow") if slow_tol is None else slow_tol(o)
    assert int(o["mine"][0]) == oracle_sample(cfg, o["my_slow"], 0, o["window"], T, p, rp, o["noise"]), f"{where}: slow-head sampler"
    assert int(o["mine"][1]) == max(int(o["mine"][0]) - cfg.semantic_begin_id, 0)
    same = int(o["mine"][0]) == int(o["ref"][0])
    for k in range(1, cfg.num_codebooks):
        if same:
            a_k = atol
            if "alt_fast" in o and bool((o["alt"][: k + 1] == o["ref"][: k + 1]).all()):
                # yardstick: the reference's own CPU path fed the same inputs up to this head -- we may be as far from its
                # CUDA path as that is (x1.25), where this exceeds the fixed tolerance
                a_k = max(atol, 1.25 * (o["alt_fast"][k - 1].float() - o["ref_fast"][k - 1].float()).abs().max().item())
            logits_close(o["my_fast"][k - 1], o["ref_fast"][k - 1], None, f"{where} fast head {k}", atol=a_k, ulps=ulps)
        assert int(o["mine"][k + 1]) == oracle_sample(cfg, o["my_fast"][k - 1], k, o["window"], T, p, rp, o["noise"]), f"{where}: fast head {k} sampler"
        same = same and int(o["mine"][k + 1]) == int(o["ref"][k + 1])
    return same


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(variant_configs().keys()))
@pytest.mark.parametrize("mode", list(MODES.keys()))
def test_tiny_teacher_forced(name, mode):
    cfg = variant_configs()[name]
    T, p, rp = MODES[mode]
    m, eng, sd = build_pair(cfg, seed=0)
    tf = TeacherForced(cfg, m, eng, synthetic_prompt(cfg, 5, 12, 4, seed=1), T, p, rp)
    agree = sum(check_step(cfg, tf.step(), T, p, rp, f"{name}/{mode} step {s}") for s in range(24))
    eng.close()
    assert agree >= 20, f"only {agree}/24 steps token-identical"


def test_long_context_multi_tile_attention():
    """6000-position prompt: more 64-row tiles than splits, so every CTA walks several double-buffered bulk-copy
    tiles and the per-head merge combines all splits"""
    cfg = tiny_config(max_seq_len=8192)
    T, p, rp = MODES["sampled"]
    m, eng, sd = build_pair(cfg, seed=0)
    tf = TeacherForced(cfg, m, eng, synthetic_prompt(cfg, 5, 5990, 5, seed=1), T, p, rp)
    agree = sum(check_step(cfg, tf.step(), T, p, rp, f"long-context step {s}") for s in range(6))
    eng.close()
    # and the engine's own prefill over the same prompt agrees with the oracle's KV state: first generated column
    m2, eng2, _ = build_pair(cfg, seed=0, bind_kv=False)
    prompt = synthetic_prompt(cfg, 5, 5990, 5, seed=1)
    noise = torch.cat([eng2.step_noise(3, s) for s in range(4)])
    eng2.set_noise(noise)
    mine = eng2.generate(prompt, 4, T, p, rp)
    logits = eng2.read("slow_logits_raw")
    eng2.close()
    assert agree >= 5 and mine.shape[1] == 4 and torch.isfinite(logits.float()).all()


@pytest.mark.parametrize("name", list(variant_configs().keys()) + ["s1mini"])
def test_mega_kernel_vs_per_phase_kernels(name):
    """the persistent whole-step kernel (mega.cuh: tensor-core dot products, chunk partials folded in order) against the
    one-kernel-per-phase path (fp32 FMA chains): same formulas and rounding points everywhere else, so after an identical
    prefill + first decode step the logits agree to accumulation-order noise and the sampled ids are the same unless the
    logits themselves are tied within that noise"""
    cfg = s1_mini_config() if name == "s1mini" else variant_configs()[name]
    sd = make_state_dict(cfg, seed=0)
    prompt = synthetic_prompt(cfg, 5, 12, 4, seed=1)
    outs = []
    for flag in (1, 0):
        eng = DualAREngine(cfg, sd, device=0, seed=5, options={"mega_kernel": flag})
        toks = eng.generate(prompt, 1, 0.7, 0.8, 1.1)
        outs.append((toks, eng.read("fast_logits").clone(), eng.read("slow_logits_raw").clone(), eng.launches_per_step()))
        eng.close()
    (ta, fa, sa, la), (tb, fb, sb, lb) = outs
    assert la[0] == 1 and lb[0] > 1
    logits_close(sa, sb, cfg, f"{name}: slow logits, persistent vs per-phase", atol=5e-2, ulps=8.0)
    if int(ta[0, 0]) == int(tb[0, 0]):
        logits_close(fa[0], fb[0], None, f"{name}: first fast head", atol=5e-2, ulps=8.0)
    else:
        assert near_tie(sb, int(ta[0, 0]), int(tb[0, 0]), ulps=8.0, atol=5e-2)


@pytest.mark.parametrize("name", ["s1like", "biased", "s1mini"])
def test_code_table_vs_wqkv_phase(name):
    """option fast_qkv_table: passes >= 1 of the fast stack take their first q | k | v from the per-code table (an fp32 FMA
    dot product per row at finalize) instead of running the wqkv phase (tensor-core chunks).  The slow stack is untouched, so
    slow logits and the semantic id are identical bit for bit; the first fast head -- the first consumer of a table row --
    agrees to accumulation-order noise."""
    cfg = s1_mini_config() if name == "s1mini" else variant_configs()[name]
    sd = make_state_dict(cfg, seed=0)
    prompt = synthetic_prompt(cfg, 5, 12, 4, seed=1)
    outs = []
    for flag in (1, 0):
        eng = DualAREngine(cfg, sd, device=0, seed=5, options={"fast_qkv_table": flag})
        toks = eng.generate(prompt, 1, 0.7, 0.8, 1.1)
        outs.append((toks, eng.read("fast_logits").clone(), eng.read("slow_logits_raw").clone()))
        eng.close()
    (ta, fa, sa), (tb, fb, sb) = outs
    assert torch.equal(sa, sb) and int(ta[0, 0]) == int(tb[0, 0]) and int(ta[1, 0]) == int(tb[1, 0])
    logits_close(fa[0], fb[0], None, f"{name}: first fast head, code table vs wqkv phase", atol=5e-2, ulps=8.0)


@pytest.fixture(scope="module")
def s1():
    cfg = s1_mini_config()
    m, eng, sd = build_pair(cfg, seed=0)
    yield cfg, m, eng, sd
    eng.close()


def fullsize_teacher_forced(cfg, m, eng, sd, n_steps, label):
    """full-size model, teacher-forced: OUR logits and torch's own bf16 CUDA logits (the oracle = the reference's ops) are both
    measured against GROUND TRUTH -- the oracle with every linear accumulated in fp64 and rounded to bf16 at the reference's
    rounding points (OracleModel.accum), stepped along the same trajectory with its own KV cache.  Any two bf16 pipelines that
    differ only in fp32 summation order decorrelate after 28 + 4 layers; the literal 2e-2 max-abs of the north star holds for
    neither.  What is asserted: we are no further from the truth than cuBLAS is (x1.15 on the mean, x1.25 + one bf16 ulp on the max,
    over the semantic logits -- the only ids that can be sampled), the mean distance to the bf16 oracle stays below 1e-2, and every
    logit lies within max(5e-2, 8 bf16 ulp) of it."""
    truth = orc.OracleModel.build(cfg, sd, device="cuda:0")
    truth.accum = torch.float64
    sem = slice(cfg.semantic_begin_id, cfg.semantic_end_id + 1)
    T, p, rp = MODES["sampled"]
    tf = TeacherForced(cfg, m, eng, synthetic_prompt(cfg, 5, 12, 4, seed=1), T, p, rp, alt=truth)
    rows = []

    def tol(o):
        mine, cuda, tr = o["my_slow"].float(), o["ref_slow"].float(), o["alt_slow"].float()
        dm, dc, d = (mine - tr).abs()[sem], (cuda - tr).abs()[sem], (mine - cuda).abs()
        rows.append((dm.max().item(), dm.mean().item(), dc.max().item(), dc.mean().item(), d[sem].max().item()))
        lim = torch.maximum(torch.full_like(d, 5e-2), 8 * bf16_ulp(o["ref_slow"]))
        assert (d <= lim).all(), f"logits beyond max(5e-2, 8 bf16 ulp) of the bf16 oracle: worst {d.max().item()}"
        assert d[sem].mean().item() < 1e-2, "mean |ours - bf16 oracle| over the semantic logits must stay below 1e-2"

    agree = sum(check_step(cfg, tf.step(), T, p, rp, f"{label} step {s}", slow_tol=tol, atol=7e-2, ulps=8.0) for s in range(n_steps))
    r = torch.tensor(rows)
    print(f"\n[{label}] semantic logits vs fp64-accumulate truth over {n_steps} teacher-forced steps:  ours max {r[:, 0].max():.4f} mean {r[:, 1].mean():.5f}   |   "
          f"torch CUDA bf16 max {r[:, 2].max():.4f} mean {r[:, 3].mean():.5f}   |   ours vs torch CUDA max {r[:, 4].max():.4f};  {agree}/{n_steps} steps token-identical")
    for i, row in enumerate(rows):
        print(f"    step {i}: ours-truth max {row[0]:.4f} mean {row[1]:.5f}; cuda-truth max {row[2]:.4f} mean {row[3]:.5f}")
    assert r[:, 1].mean().item() <= 1.15 * r[:, 3].mean().item() + 1e-4, "mean distance to the truth exceeds cuBLAS's"
    assert r[:, 0].max().item() <= 1.25 * r[:, 2].max().item() + 2.0 ** -6, "max distance to the truth exceeds cuBLAS's"
    return agree


def test_s1_mini_teacher_forced_with_yardstick(s1):
    cfg, m, eng, sd = s1
    assert fullsize_teacher_forced(cfg, m, eng, sd, 8, "s1-mini") >= 6


def test_fish_speech_1_5_shape_teacher_forced():
    """the second BASELINE shape at full size (24 layers, 2 kv heads of 64, untied 102,048-row head, 8 codebooks of 1024) against the
    oracle and the fp64-accumulate truth -- not only against itself"""
    from fish_tts_b200.config import fish_speech_1_5_config
    cfg = fish_speech_1_5_config()
    m, eng, sd = build_pair(cfg, seed=0)
    try:
        assert fullsize_teacher_forced(cfg, m, eng, sd, 6, "fish-speech-1.5 shape") >= 4
    finally:
        eng.close()


def free_running(cfg, m, eng, prompt, n, T, p, rp, seed):
    """engine and oracle each run FREE (own trajectory, own KV cache) under the same Philox noise; returns the engine's columns with
    its per-step raw logits (slow, fast), the oracle's columns with its per-step traces, and the noise"""
    noise = torch.cat([eng.step_noise(seed, s) for s in range(n)])
    eng.set_noise(noise)
    eng.prefill(prompt, n, T, p, rp)
    my_slow, my_fast = [], []
    for s in range(n):
        if s > 0:
            eng.decode(1)
        torch.cuda.synchronize()
        my_slow.append(eng.read("slow_logits_raw").clone()); my_fast.append(eng.read("fast_logits").clone())
    mine, _ = eng.collect()
    eng.set_noise(None)
    tr = []

    def noise_fn(call, k):      # slow head, then the fast heads of each step, in the oracle's sampling order
        step, head = divmod(call, cfg.num_codebooks)
        off = step * eng.noise_per_step + (0 if head == 0 else cfg.vocab_size + (head - 1) * eng.fast_vocab)
        return noise[off: off + k]

    with torch.inference_mode(), sdpa_kernel(SDPBackend.MATH):
        ref = orc.generate(m, prompt.to(m.device), n, T, p, rp, noise=orc.NoiseSource(noise_fn), stable_ties=True, trace=tr, math_prefill=True)
    return mine, my_slow, my_fast, ref[:, prompt.size(1):].cpu().numpy(), tr, noise


@pytest.mark.parametrize("mode", ["greedy", "sampled"])
def test_s1_mini_free_running_streams(mode):
    """north star: 'greedy token ids bit-exact for the first 256 steps; sampled streams identical under a shared seeded Philox RNG'.
    Both run FREE here (no teacher forcing, own KV caches).  Two bf16 pipelines with different fp32 summation orders part ways at the
    first decision whose margin is below the logit noise (measured against the fp64-accumulate truth in the yardstick test: max
    ~0.05, the same for cuBLAS); from there on the inputs differ and a comparison means nothing.  Reported: the first divergent
    (step, head) and what the ORACLE sees there.  Asserted: everything before it is bit-identical; AT it our logits are within the
    per-logit tolerance of the oracle's (the trajectories are still identical), our id is exactly what the reference sampler draws
    from OUR logits with the shared noise, and -- greedy -- the oracle scores the two ids within twice the tolerance."""
    cfg = s1_mini_config()
    m, eng, sd = build_pair(cfg, seed=0, bind_kv=False)
    T, p, rp = MODES[mode]
    n = 256 if mode == "greedy" else 96
    prompt = synthetic_prompt(cfg, 5, 12, 4, seed=1)
    try:
        mine, my_slow, my_fast, ref, tr, noise = free_running(cfg, m, eng, prompt, n, T, p, rp, seed=77)
        per_step = eng.noise_per_step
    finally:
        eng.close()
    assert mine.shape == ref.shape == (cfg.num_codebooks + 1, n)
    same = (mine == ref).all(axis=0)
    first = int(np.argmin(same)) if not same.all() else n
    msg = f"\n[s1-mini free-running {mode}] {n} steps: all {cfg.num_codebooks + 1} rows bit-identical for the first {first} steps"
    if first < n:
        row = int(np.argmin(mine[:, first] == ref[:, first]))          # first differing row: 0 = semantic id, k >= 2 = codebook k-1 (fast head k-1)
        head = 0 if row == 0 else row - 1
        assert row != 1, "codebook 0 is a function of the semantic id"
        lg_ref = (tr[first].slow_logits if head == 0 else tr[first].fast_logits[head - 1]).float().cpu()
        lg_mine = (my_slow[first] if head == 0 else my_fast[first][head - 1]).float()
        a, b = int(mine[row, first]), int(ref[row, first])
        ia, ib = (a, b)
        d = (lg_mine - lg_ref).abs()
        msg += (f"; first divergence at step {first}, head {head}: oracle id {ib} (its logit {lg_ref[ib]:.4f}), ours {ia} (oracle's logit for it {lg_ref[ia]:.4f}, "
                f"margin {float(lg_ref[ib] - lg_ref[ia]):.4f}); |our logits - oracle logits| at that head: max {d.max():.4f}")
        print(msg)
        lim = torch.maximum(torch.full_like(d, 7e-2 if head else 5e-2), 8 * bf16_ulp(lg_ref))
        assert (d <= lim).all(), "logits at the first divergent head are outside the per-logit tolerance although the inputs are identical"
        # our id is exactly the reference sampler's draw from OUR logits (window: the engine's own history, identical to the oracle's so far)
        i = first - 1
        prev = np.zeros((cfg.num_codebooks + 1, cfg.max_seq_len), dtype=np.int32)
        if first > 1:
            prev[:, : first - 1] = mine[:, 1:first]
        window = None if first == 0 else torch.from_numpy(prev[:, :16] if i < 16 else prev[:, i - 16: i])
        blk = noise[first * per_step: (first + 1) * per_step]
        raw = my_slow[first] if head == 0 else my_fast[first][head - 1]
        assert a == oracle_sample(cfg, raw, head, window, T, p, rp, blk), "our id is not the reference sampler's draw from our logits"
        if mode == "greedy":
            tol = max(7e-2 if head else 5e-2, 8 * float(bf16_ulp(lg_ref[ib])))
            assert float(lg_ref[ib] - lg_ref[ia]) <= 2 * tol, f"decisive greedy divergence: margin {float(lg_ref[ib] - lg_ref[ia]):.4f}"
    else:
        print(msg)
    assert first >= 1, "the very first token already differs"


# ---- 1. the sampler alone -------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def s1_per_phase():
    cfg = s1_mini_config()
    sd = make_state_dict(cfg, seed=0)
    eng = DualAREngine(cfg, sd, device=0, options={"mega_kernel": 0})
    yield cfg, eng
    eng.close()


def adversarial_logits(dist, n, cfg, head, g):
    if dist == "peaky":
        lg = torch.randn(n, generator=g) * 4.0
    elif dist == "conditioned":
        lg = torch.randn(n, generator=g) * 0.64
        if head == 0:
            lg[: cfg.semantic_begin_id] -= 8.0
            lg[cfg.semantic_end_id + 1:] -= 8.0
    elif dist == "flat":
        lg = torch.randn(n, generator=g) * 0.3          # every logit is a candidate (> DA_CAND_CAP): whole-vocabulary fallback walk
    elif dist == "ties":
        lg = torch.randint(-3, 4, (n,), generator=g).float() * 0.5
    elif dist == "mid6000":
        # 4096 < candidates <= DA_CAND_CAP (8192): the persistent kernel's binned sampler declines (fallback), the per-phase
        # kernels' bisection sampler still takes the list
        lg = torch.full((n,), -9.0)
        k = min(6000, n // 2)
        lg[torch.randperm(n, generator=g)[:k]] = torch.randn(k, generator=g) * 0.5
    elif dist == "plateau":
        # few candidates, but the nucleus wants more than the list holds (mass just below the candidate threshold): the list
        # samplers must notice and hand over to the whole-vocabulary walk
        lg = torch.full((n,), -6.5) + torch.randn(n, generator=g) * 0.05
        k = min(300, n // 4)
        lg[torch.randperm(n, generator=g)[:k]] = torch.randn(k, generator=g) * 0.1
    else:
        lg = torch.full((n,), -5.0); lg[int(torch.randint(0, n, (1,), generator=g))] = 9.0
    return lg.bfloat16()


@pytest.mark.parametrize("path", ["persistent_kernel", "per_phase_kernels"])
@pytest.mark.parametrize("dist", ["peaky", "conditioned", "flat", "ties", "one_hot", "mid6000", "plateau"])
def test_sampler_exact_on_identical_logits(s1, s1_per_phase, path, dist):
    """the samplers of BOTH kernel paths against the reference sampler on identical adversarial logits, bit-exact.
    persistent_kernel: dualar_debug_sample runs one whole step of mega_kernel with its logits epilogues fed these logits, i.e. the
    product path's penalty + per-CTA statistics + ordered candidate list + sample_binned / sample_sorted / sample_fallback<CBlock>
    (slow head) and the 16-warp binned sampler (fast heads)."""
    cfg = s1[0]
    eng = s1[2] if path == "persistent_kernel" else s1_per_phase[1]
    g = torch.Generator().manual_seed({"peaky": 1, "conditioned": 2, "flat": 3, "ties": 4, "one_hot": 5, "mid6000": 6, "plateau": 7}[dist])
    V, fv = cfg.vocab_size, 1024
    for trial, (T, p, rp) in enumerate([(0.7, 0.8, 1.1), (0.7, 0.7, 1.5), (1.0, 1.0, 1.0), (0.7, 1e-9, 1.0), (0.3, 0.5, 1.2), (1.5, 0.95, 0.8)]):
        for head in (0, 3):
            n = V if head == 0 else fv
            lg = adversarial_logits(dist, n, cfg, head, g)
            window = torch.randint(0, fv, (cfg.num_codebooks + 1, 16), generator=g, dtype=torch.int32).cuda()
            if head == 0 and dist in ("conditioned", "mid6000"):      # put the penalised ids where the mass is
                window[:, 0] = torch.topk(lg.float(), cfg.num_codebooks + 1).indices.to(torch.int32).cuda()
            blk = eng.step_noise(5, trial)
            t = [torch.tensor(v, device="cuda", dtype=torch.float) for v in (T, p, rp)]
            mine = eng.debug_sample(head, lg, window, *t, blk)
            ref = oracle_sample(cfg, lg, head, window, T, p, rp, blk)
            assert mine == ref, f"{path} {dist} head {head} T={T} p={p} rp={rp}: {mine} != {ref}"


# ---- 3. the engine-owned loop ----------------------------------------------------------------------------------
def run_loop_against_oracle(cfg, m, eng, prompt, n_new, T, p, rp, noise_seed):
    """Drive the ENGINE's loop (own prefill, own KV cache, device-side window / position / EOS / noise bookkeeping) one
    step at a time and replay ITS trajectory through the oracle: at every step the oracle, fed the engine's previous
    column and the window the reference would build from the engine's history (inference.py:186-191), must produce
    logits within tolerance, and the engine's ids must be what the reference sampler draws from the engine's logits.
    Returns the engine's columns (C+1, n)."""
    dev, C1 = m.device, cfg.num_codebooks + 1
    t = [torch.tensor(v, device=dev, dtype=torch.float) for v in (T, p, rp)]
    noise = torch.cat([eng.step_noise(noise_seed, s) for s in range(n_new)])
    eng.set_noise(noise)
    eng.prefill(prompt, n_new, T, p, rp)
    m.setup_caches(cfg.max_seq_len)
    pr, Tlen = prompt.to(dev), prompt.size(1)
    prev = torch.zeros((C1, cfg.max_seq_len), dtype=torch.int32, device=dev)
    cols = None
    for s in range(n_new):
        if s > 0:
            eng.decode(1)
        cols, fin = eng.collect()
        if cols.shape[1] <= s:
            assert fin
            break
        my_slow, my_fast = eng.read("slow_logits_raw"), eng.read("fast_logits")
        blk = noise[s * eng.noise_per_step: (s + 1) * eng.noise_per_step]
        tr = []
        with torch.inference_mode(), sdpa_kernel(SDPBackend.MATH):
            if s == 0:
                for i in range(Tlen - 1):
                    orc.forward_generate(m, pr[:, i:i + 1].view(1, C1, 1), torch.tensor([i], device=dev))
                window = None
                ref = orc.decode_one_token_ar(m, pr[:, -1:].view(1, C1, 1), torch.tensor([Tlen - 1], device=dev), *t, None,
                                              noise=block_noise_source(cfg, blk), stable_ties=True, trace=tr)
            else:
                i = s - 1                                   # iteration index of decode_n_tokens
                window = prev[:, :16] if i < 16 else prev[:, i - 16: i]
                cur = torch.from_numpy(cols[:, s - 1]).to(dev).view(1, C1, 1)
                ref = orc.decode_one_token_ar(m, cur, torch.tensor([Tlen + i], device=dev, dtype=torch.int32), *t, window,
                                              noise=block_noise_source(cfg, blk), stable_ties=True, trace=tr)
        o = {"mine": torch.from_numpy(cols[:, s]), "ref": ref[:, 0].cpu(), "my_slow": my_slow, "my_fast": my_fast,
             "ref_slow": tr[0].slow_logits.cpu(), "ref_fast": torch.stack(tr[0].fast_logits).cpu(),
             "window": None if window is None else window.clone().cpu(), "noise": blk}
        check_step(cfg, o, T, p, rp, f"loop step {s}")
        if s > 0:
            prev[:, s - 1] = torch.from_numpy(cols[:, s]).to(dev)      # previous_tokens[:, i] = column i+1
        # the reference's stop rule: <|im_end|> is tested on columns produced inside decode_n_tokens only
        should_stop = (s > 0 and int(cols[0, s]) == cfg.im_end_id) or s == n_new - 1
        assert fin == should_stop, f"step {s}: finished={fin}, reference rule says {should_stop}"
        if fin:
            break
    return cols


@pytest.mark.parametrize("name", list(variant_configs().keys()))
@pytest.mark.parametrize("mode", ["sampled", "greedy"])
def test_loop_machinery_against_oracle(name, mode):
    cfg = variant_configs()[name]
    T, p, rp = MODES[mode]
    m, eng, sd = build_pair(cfg, seed=0, bind_kv=False)
    prompt = synthetic_prompt(cfg, 5, 12, 4, seed=1)
    cols = run_loop_against_oracle(cfg, m, eng, prompt, 40, T, p, rp, noise_seed=21)
    # one-shot generate() reproduces the stepwise run bit for bit (same noise, same engine)
    again = eng.generate(prompt, 40, T, p, rp)
    eng.close()
    assert cols.shape == (cfg.num_codebooks + 1, 40) and (again == cols).all()


@pytest.mark.parametrize("name", list(variant_configs().keys()))
def test_eos_stops_the_loop(name):
    """<|im_end|> reachable: the device-side flag ends the request after recording the column (inference.py:208-211)"""
    cfg = variant_configs()[name]
    m, eng, sd = build_pair(cfg, seed=0, bind_kv=False, eos_reachable=True)
    prompt = synthetic_prompt(cfg, 5, 12, 4, seed=1)
    T, p, rp = MODES["sampled"]
    stopped = 0
    for trial in range(4):
        out = run_loop_against_oracle(cfg, m, eng, prompt, 24, T, p, rp, noise_seed=100 + trial)
        if out.shape[1] < 24:
            stopped += 1
            assert out[0, -1] == cfg.im_end_id and (out[0, 1:-1] != cfg.im_end_id).all()
            eng.decode(5)                                   # replays after the end are no-ops
            again, fin = eng.collect()
            assert fin and again.shape == out.shape and (again == out).all()
    eng.close()
    assert stopped >= 1, "the eos-reachable checkpoint never produced <|im_end|>"


# ---- 4. fixtures of the unmodified reference (CPU) ---------------------------------------------------------------
@pytest.mark.parametrize("name", list(variant_configs().keys()))
@pytest.mark.parametrize("mode", list(MODES.keys()))
def test_golden_reference_fixtures(name, mode):
    """teacher-forced along the reference's own CPU trajectory, logits against its recorded logits"""
    cfg = variant_configs()[name]
    g = torch.load(GOLD / f"tiny_{name}_{mode}.pt")
    sd = make_state_dict(cfg, seed=0)
    eng = DualAREngine(cfg, sd, device=0, options={"cpu_scalar_semantics": 1})
    prompt, Tlen = g["prompt"], g["prompt"].size(1)
    n_steps = g["tokens"].size(0)
    noise = torch.cat([philox.step_noise(cfg, g["noise_seed"], s) for s in range(n_steps)]).cuda()
    eng.set_noise(noise)
    eng.prefill(prompt, 2, g["temperature"], g["top_p"], g["repetition_penalty"])   # fills the KV cache, produces column 0
    first, _ = eng.collect()
    logits_close(eng.read("slow_logits_raw"), g["slow_logits"][0], cfg, f"{name}/{mode} prefill logits")
    t = [torch.tensor(v, device="cuda", dtype=torch.float) for v in (g["temperature"], g["top_p"], g["repetition_penalty"])]
    prev = torch.zeros((cfg.num_codebooks + 1, cfg.max_seq_len), dtype=torch.int32, device="cuda")
    same = int((first[:, 0] == g["tokens"][0].numpy()).all())
    for i in range(n_steps - 1):
        cur = g["tokens"][i].cuda().view(1, -1, 1)
        window = prev[:, :16] if i < 16 else prev[:, i - 16: i]
        blk = noise[(i + 1) * eng.noise_per_step: (i + 2) * eng.noise_per_step]
        mine = eng.step(cur, torch.tensor([Tlen + i], device="cuda", dtype=torch.int32), window, *t, noise=blk).clone().cpu()
        logits_close(eng.read("slow_logits_raw"), g["slow_logits"][i + 1], cfg, f"{name}/{mode} step {i}")
        ok = bool((mine[:, 0] == g["tokens"][i + 1]).all())
        if ok:
            logits_close(eng.read("fast_logits"), g["fast_logits"][i + 1], None, f"{name}/{mode} step {i} fast")
        same += ok
        prev[:, i] = g["tokens"][i + 1].cuda()
    eng.close()
    assert same >= int(0.8 * n_steps), f"{same}/{n_steps} steps reproduce the reference's ids"


def test_golden_s1_mini_fixture(s1):
    """the full-size fixture of the unmodified reference: top-256 and strided logits of the first steps"""
    cfg, m, eng, sd = s1
    g = torch.load(GOLD / "s1mini_sampled.pt")
    eng2 = DualAREngine(cfg, sd, device=0, options={"cpu_scalar_semantics": 1})
    noise = torch.cat([philox.step_noise(cfg, g["noise_seed"], s) for s in range(g["tokens"].size(0))]).cuda()
    eng2.set_noise(noise)
    eng2.prefill(g["prompt"], 2, g["temperature"], g["top_p"], g["repetition_penalty"])
    eng2.collect()
    mine = eng2.read("slow_logits_raw").float()
    ref_top, idx = g["slow_top_val"][0].float(), g["slow_top_idx"][0].long()
    d = (mine[idx] - ref_top).abs()
    assert d.max().item() <= 0.05 and d.mean().item() < 1e-2, (d.max().item(), d.mean().item())
    ds = (mine[::16] - g["slow_strided"][0].float()).abs()
    assert (ds <= torch.maximum(torch.full_like(ds, 5e-2), 8 * bf16_ulp(g["slow_strided"][0]))).all(), ds.max().item()
    # the non-semantic logits (~ -28, ulp 0.125) are dominated by one direction of the hidden state and move TOGETHER by a few
    # ulp between any two accumulation orders (tensor-core vs fp32 FMA chains); they are never sampled
    assert (ds / bf16_ulp(g["slow_strided"][0])).mean().item() < 4.0
    eng2.close()


# ---- 5. properties at BASELINE sizes ---------------------------------------------------------------------------------
def test_full_size_generation_properties(s1):
    cfg, m, eng_bound, sd = s1
    eng = DualAREngine(cfg, sd, device=0, seed=99)
    prompt = synthetic_prompt(cfg, 3, 215, 5, seed=1)            # BASELINE configs[1]: ~220-token prompt
    a = eng.generate(prompt, 1024, 0.7, 0.8, 1.1)
    assert a.shape == (11, 1024), "EOS must be unreachable on the conditioned checkpoint"
    assert ((a[0] >= cfg.semantic_begin_id) & (a[0] <= cfg.semantic_end_id)).all()
    assert (a[1] == a[0] - cfg.semantic_begin_id).all()           # inference.py:123-126
    assert (a[2:] >= 0).all() and (a[2:] < 1024).all()
    eng.seed(99)
    b = eng.generate(prompt, 1024, 0.7, 0.8, 1.1)
    assert (a == b).all(), "same seed, same request => identical stream (idempotent replay)"
    eng.seed(100)
    c = eng.generate(prompt, 64, 0.7, 0.8, 1.1)
    assert not (c == a[:, :64]).all()
    # clamp to the cache: inference.py:301-307
    d = eng.generate(synthetic_prompt(cfg, 3, cfg.max_seq_len - 16 - 8, 5, seed=2), 0, 0.7, 0.8, 1.1)   # T = max_seq_len - 16
    assert d.shape[1] == 16
    eng.close()


def test_noise_kernel_matches_host_philox(s1):
    cfg, m, eng, sd = s1
    for (seed, step, head, n) in [(0, 0, 0, 155776), (1234567890123, 7, 3, 1024), (2 ** 63 + 5, 1000, 9, 4096)]:
        dev = eng.fill_noise(seed, step, head, n).cpu().float()
        host = philox.exp1_noise(seed, step, head, n).float()
        neq = (dev != host)
        assert neq.float().mean().item() < 1e-3, "same Philox counters; only logf rounding may differ"
        assert ((dev - host).abs() <= bf16_ulp(host)).all()


def test_api_errors():
    cfg = tiny_config()
    sd = make_state_dict(cfg, seed=0)
    with pytest.raises(capi.DualarError, match="never loaded"):
        DualAREngine(cfg, {k: v for k, v in sd.items() if k != "norm.weight"}, device=0)
    with pytest.raises(capi.DualarError, match="unknown weight"):
        DualAREngine(cfg, dict(sd, **{"fast_project_in.weight": torch.zeros(4)}), device=0)
    eng = DualAREngine(cfg, sd, device=0)
    with pytest.raises(capi.DualarError, match="without a prefilled request"):
        eng.decode(1)
    with pytest.raises(capi.DualarError, match="exceeds max_seq_len"):
        eng.generate(np.zeros((cfg.num_codebooks + 1, cfg.max_seq_len), dtype=np.int32), 4)
    out = eng.generate(synthetic_prompt(cfg, 2, 3, 1), 5)      # still usable after errors
    assert out.shape[1] == 5
    eng.close()


def test_fish_speech_1_5_shape_full_size():
    """BASELINE configs[2] shape (24 layers, GQA 8, untied 102k head, 8 x 1024 codebooks) at full size on the persistent
    kernel: id ranges, determinism under a fixed seed, and agreement of the first decode step with the per-phase kernels"""
    from fish_tts_b200.config import fish_speech_1_5_config
    cfg = fish_speech_1_5_config()
    sd = make_state_dict(cfg, seed=0)
    prompt = synthetic_prompt(cfg, 5, 12, 4, seed=1)
    runs = []
    for flag in (1, 1, 0):
        eng = DualAREngine(cfg, sd, device=0, seed=5, options={"mega_kernel": flag})
        toks = torch.as_tensor(eng.generate(prompt, 24 if flag else 1, 0.7, 0.8, 1.1))
        runs.append((toks, eng.read("slow_logits_raw").clone(), eng.launches_per_step()))
        eng.close()
    (t1, l1, n1), (t2, l2, n2), (t0, l0, n0) = runs
    assert n1[0] == 1 and n0[0] > 1, "the 1.5 shape must run on the persistent kernel"
    assert torch.equal(t1, t2) and torch.equal(l1, l2), "same seed, same tokens and logits"
    assert ((t1[0] >= cfg.semantic_begin_id) & (t1[0] <= cfg.semantic_end_id)).all()
    assert ((t1[1:] >= 0) & (t1[1:] < cfg.codebook_size)).all()
    assert int(t1[0, 0]) == int(t0[0, 0]) or near_tie(l0, int(t1[0, 0]), int(t0[0, 0]), ulps=8.0, atol=5e-2)


def test_binned_sampler_equals_sorting_sampler():
    """The slow head's binned nucleus sampler against the sorting one, in isolation (tests/cuda/binned_check.cu): 1152 random
    logit vectors (flat, peaked, heavy ties, all equal, long tails, a far-away half; top_p from 1e-9 to 1; items = the whole
    vocabulary or candidates of a larger one) -- token and nucleus size must agree in every case."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    root = Path(__file__).resolve().parent.parent
    exe = root / "build" / "binned_check"
    src = root / "tests" / "cuda" / "binned_check.cu"
    deps = [src] + list((root / "fish_tts_b200" / "csrc").glob("*.cuh"))
    if not exe.exists() or any(d.stat().st_mtime > exe.stat().st_mtime for d in deps):
        exe.parent.mkdir(exist_ok=True)
        subprocess.run([nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe), str(src)], check=True, timeout=600)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatches 0" in r.stdout, r.stdout[-2000:]
