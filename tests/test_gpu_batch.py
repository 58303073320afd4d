"""Parity of the many-column path (tcgen05 GEMMs: tensor-core prefill + batched decode) -- runs on the B200 box.

The reference is batch 1 only (inference.py:73 reads ``logits[0, -1]``; :355 views the prompt as (1, C+1, T)), so the
oracle for a batch of B requests is B independent runs of the batch-1 oracle (SURVEY.md section 7 "hard parts", 8d config 4):

  * every slot's trajectory is replayed through the oracle (prefill of the whole prompt in ONE forward like
    inference.py:353-362, then step by step on the ENGINE's tokens): per-step logits within the tolerance of the batch-1
    tests, and every id the engine drew is what the reference sampler draws from the engine's logits and noise;
  * batch invariance: a request's tokens and logits are bit-identical whether it runs alone or next to other requests;
  * continuous batching: releasing and refilling one slot leaves the other slots' streams untouched;
  * the tensor-core prefill writes the KV rows the oracle's one-shot prefill writes (within bf16 accumulation noise).
"""
import numpy as np
import pytest
import torch
from torch.nn.attention import SDPBackend, sdpa_kernel

from fish_tts_b200.config import s1_mini_config, tiny_config
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt
from helpers import bf16_ulp, logits_close, near_tie, variant_configs

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from fish_tts_b200.engine import DualAREngine
    from gpu_common import block_noise_source, build_pair
    from test_gpu_parity import check_step
from oracle import dualar_oracle as orc

REQS = [  # (text tokens, reference frames, seed, temperature, top_p, repetition_penalty)
    (5, 12, 1, 0.7, 0.8, 1.1), (3, 30, 2, 0.7, 1e-9, 1.0), (7, 3, 3, 1.0, 1.0, 1.5), (4, 50, 4, 0.3, 0.5, 1.2), (6, 20, 5, 0.7, 0.7, 1.5),
]


def make_prompt(cfg, req):
    return synthetic_prompt(cfg, req[0], req[1], 4, seed=req[2])


def replay_slot(cfg, m, eng, prompt, cols, slows, fasts, T, p, rp, seed, where, **tol):
    """oracle along the ENGINE's trajectory of one slot; returns the number of steps with all rows identical"""
    dev, C1 = m.device, cfg.num_codebooks + 1
    t = [torch.tensor(v, device=dev, dtype=torch.float) for v in (T, p, rp)]
    m.setup_caches(cfg.max_seq_len)
    pr, Tlen = prompt.to(dev), prompt.size(1)
    prev = torch.zeros((C1, cfg.max_seq_len), dtype=torch.int32, device=dev)
    same = 0
    for s in range(cols.shape[1]):
        blk = eng.step_noise(seed, s)
        tr = []
        with torch.inference_mode(), sdpa_kernel(SDPBackend.MATH):
            if s == 0:      # the reference's prefill: the whole prompt in one forward, no repetition penalty
                window = None
                ref = orc.decode_one_token_ar(m, pr.view(1, C1, -1), torch.arange(Tlen, device=dev), *t, None,
                                              noise=block_noise_source(cfg, blk), stable_ties=True, trace=tr)
            else:
                i = s - 1
                window = prev[:, :16] if i < 16 else prev[:, i - 16: i]
                cur = torch.from_numpy(cols[:, s - 1]).to(dev).view(1, C1, 1)
                ref = orc.decode_one_token_ar(m, cur, torch.tensor([Tlen + i], device=dev, dtype=torch.int32), *t, window,
                                              noise=block_noise_source(cfg, blk), stable_ties=True, trace=tr)
        o = {"mine": torch.from_numpy(cols[:, s]), "ref": ref[:, 0].cpu(), "my_slow": slows[s], "my_fast": fasts[s],
             "ref_slow": tr[0].slow_logits.cpu(), "ref_fast": torch.stack(tr[0].fast_logits).cpu(),
             "window": None if window is None else window.clone().cpu(), "noise": blk}
        same += bool(check_step(cfg, o, T, p, rp, f"{where} step {s}", **tol))
        if s > 0:
            prev[:, s - 1] = torch.from_numpy(cols[:, s]).to(dev)
    return same


def run_batch(eng, cfg, reqs, n_steps, slots=None, seeds=None):
    """prefill every request into its slot, then n_steps batched steps; per slot: columns, per-step raw logits"""
    slots = slots or list(range(len(reqs)))
    prompts = [make_prompt(cfg, r) for r in reqs]
    for sl, r, pr in zip(slots, reqs, prompts):
        eng.batch_prefill(sl, pr, n_steps, r[3], r[4], r[5], seed=(seeds[sl] if seeds else 100 + sl))
    slows = {sl: [] for sl in slots}
    fasts = {sl: [] for sl in slots}
    for s in range(n_steps):
        eng.batch_decode(1)
        sl_all, fa_all = eng.batch_read("slow_logits_raw"), eng.batch_read("fast_logits")
        for sl in slots:
            slows[sl].append(sl_all[sl].clone())
            fasts[sl].append(fa_all[sl].clone())
    cols = {sl: eng.batch_collect(sl)[0] for sl in slots}
    return prompts, cols, slows, fasts


@pytest.mark.parametrize("name", list(variant_configs().keys()))
def test_batched_decode_equals_independent_oracle_runs(name):
    cfg = variant_configs()[name]
    sd = make_state_dict(cfg, seed=0)
    m = orc.OracleModel.build(cfg, sd, device="cuda:0")
    eng = DualAREngine(cfg, sd, device=0)
    eng.batch_init(5, cfg.max_seq_len)
    n = 12
    prompts, cols, slows, fasts = run_batch(eng, cfg, REQS, n)
    total_same = 0
    for sl, r in enumerate(REQS):
        assert cols[sl].shape == (cfg.num_codebooks + 1, n)
        total_same += replay_slot(cfg, m, eng, prompts[sl], cols[sl], slows[sl], fasts[sl], r[3], r[4], r[5], 100 + sl, f"{name} slot {sl}")
    launches = int(eng.batch_read("launches")[0])
    eng.close()
    print(f"\n[{name}] batched step = {launches} kernels; {total_same}/{n * len(REQS)} (slot, step) pairs identical to the oracle in all rows")
    assert total_same >= int(0.8 * n * len(REQS))


def test_batch_invariance_and_slot_independence():
    """a request alone in the batch == the same request next to four others, bit for bit (tokens AND logits); and the slot it
    sits in does not matter"""
    cfg = variant_configs()["s1like"]
    sd = make_state_dict(cfg, seed=0)
    eng = DualAREngine(cfg, sd, device=0)
    eng.batch_init(5, cfg.max_seq_len)
    n = 10
    _, cols_a, slows_a, fasts_a = run_batch(eng, cfg, REQS, n)
    for sl in range(5):
        eng.batch_release(sl)
    # request 2 alone, in slot 4, with the seed it had in slot 2
    _, cols_b, slows_b, fasts_b = run_batch(eng, cfg, [REQS[2]], n, slots=[4], seeds={4: 102})
    eng.close()
    assert (cols_a[2] == cols_b[4]).all()
    for s in range(n):
        assert torch.equal(slows_a[2][s], slows_b[4][s]) and torch.equal(fasts_a[2][s], fasts_b[4][s]), f"step {s}: logits depend on the batch"


def test_fused_norm_gemm_equals_separate_norm_kernel(monkeypatch):
    """DUALAR_TC_FUSE_NORM=1: the decode GEMMs normalise their own operand (gemm_tc_kernel<32, true>: RMSNorm applied while staging
    the activation tile in shared memory); default: a separate RMSNorm kernel runs in front of a TMA-fed GEMM.  Same formulas, same
    summation order: tokens and logits must agree bit for bit."""
    cfg = variant_configs()["s1like"]
    sd = make_state_dict(cfg, seed=0)
    outs = []
    for fuse in ("1", "0"):
        monkeypatch.setenv("DUALAR_TC_FUSE_NORM", fuse)
        eng = DualAREngine(cfg, sd, device=0)
        eng.batch_init(5, cfg.max_seq_len)
        _, cols, slows, fasts = run_batch(eng, cfg, REQS, 8)
        outs.append((cols, slows, fasts, int(eng.batch_read("launches")[0])))
        eng.close()
    (ca, sa, fa, la), (cb, sb, fb, lb) = outs
    assert la < lb, f"fusing the norms must remove launches ({la} vs {lb})"
    for sl in range(5):
        assert (ca[sl] == cb[sl]).all()
        for s in range(8):
            assert torch.equal(sa[sl][s], sb[sl][s]) and torch.equal(fa[sl][s], fb[sl][s]), f"slot {sl} step {s}"


def test_request_groups_do_not_change_a_request():
    """Slots organised in concurrently running groups (own buffers, KV, split-K workspace, graph and stream per group): every request's
    tokens and logits are bit-identical to the same request in one big group -- and a refill in one group leaves the others alone."""
    cfg = variant_configs()["s1like"]
    sd = make_state_dict(cfg, seed=0)
    outs = []
    for gs in (32, 2):
        eng = DualAREngine(cfg, sd, device=0)
        eng.batch_init(5, cfg.max_seq_len, group_slots=gs)
        assert int(eng.batch_read("groups")[0]) == (1 if gs == 32 else 3)
        _, cols, slows, fasts = run_batch(eng, cfg, REQS, 10)
        if gs == 2:      # continuous batching across groups: slot 3 (group 1) is refilled, slots in groups 0 and 2 keep going
            eng.batch_release(3)
            r = REQS[1]
            eng.batch_prefill(3, make_prompt(cfg, r), 4, r[3], r[4], r[5], seed=901)
            eng.batch_decode(4)
            again, fin = eng.batch_collect(3)
            assert fin and again.shape[1] == 4
        outs.append((cols, slows, fasts))
        eng.close()
    (ca, sa, fa), (cb, sb, fb) = outs
    for sl in range(5):
        assert (ca[sl] == cb[sl]).all(), f"slot {sl}: tokens depend on the grouping"
        for s in range(10):
            assert torch.equal(sa[sl][s], sb[sl][s]) and torch.equal(fa[sl][s], fb[sl][s]), f"slot {sl} step {s}: logits depend on the grouping"


@pytest.mark.parametrize("name", ["s1like", "projected"])
def test_persistent_step_equals_per_kernel_graph(name):
    """option batch_persistent: the batched step as two cooperative launches of bstep_kernel (one CTA per SM walks the phase table,
    grid barriers between phases, a producer warp prefetching the next GEMM's weight tiles) instead of ~100 kernels per step.  Same
    device code, same split-K partition: identical bits."""
    cfg = variant_configs()[name]
    sd = make_state_dict(cfg, seed=0)
    eng = DualAREngine(cfg, sd, device=0)
    eng.batch_init(5, cfg.max_seq_len)
    outs = []
    for persist in (0, 1):
        eng.set_option("batch_persistent", persist)
        _, cols, slows, fasts = run_batch(eng, cfg, REQS, 8)
        outs.append((cols, slows, fasts, int(eng.batch_read("launches")[0])))
        for sl in range(5):
            eng.batch_release(sl)
    eng.close()
    (ca, sa, fa, la), (cb, sb, fb, lb) = outs
    assert lb == 4 and la > lb, f"persistent step: 2 cooperative launches + the slow sampler's 2 kernels ({lb}), per-kernel graph {la}"
    for sl in range(5):
        assert (ca[sl] == cb[sl]).all()
        for s in range(8):
            assert torch.equal(sa[sl][s], sb[sl][s]) and torch.equal(fa[sl][s], fb[sl][s]), f"slot {sl} step {s}"


def test_continuous_batching_over_groups_equals_one_request_at_a_time():
    """The whole serving loop under churn: 36 short utterances through `replicas.run_rank_batched` on 8 slots in 4 groups of 2
    (asynchronous prefills beside the decoding groups, releases without a device synchronisation, slots refilled many times) --
    every utterance's codes must be those of the same request run alone in a one-slot engine."""
    from fish_tts_b200 import replicas
    cfg = variant_configs()["s1like"]
    sd = make_state_dict(cfg, seed=0)
    rng = np.random.default_rng(11)
    utts = []
    for i in range(36):
        T = int(rng.integers(10, 60))
        utts.append(replicas.Utterance(uid=i, prompt=synthetic_prompt(cfg, 3, T - 7, 4, seed=500 + i).numpy(), max_new_tokens=int(rng.integers(3, 40)),
                                       temperature=0.7, top_p=0.8, repetition_penalty=1.1))
    eng = DualAREngine(cfg, sd, device=0)
    eng.batch_init(8, cfg.max_seq_len, group_slots=2)
    assert int(eng.batch_read("groups")[0]) == 4
    res = replicas.run_rank_batched(eng, utts, 0, 1, 8, poll_steps=5, sync=torch.cuda.synchronize, seed_base=7000)
    eng.close()
    assert sorted(res.uids) == list(range(36))
    alone = DualAREngine(cfg, sd, device=0)
    alone.batch_init(1, cfg.max_seq_len)
    for u in utts:
        alone.batch_prefill(0, u.prompt, u.max_new_tokens, u.temperature, u.top_p, u.repetition_penalty, seed=7000 + u.uid)
        alone.batch_decode(u.max_new_tokens + 1)
        ref, fin = alone.batch_collect(0)
        alone.batch_release(0)
        assert fin and ref.shape == res.codes[u.uid].shape, f"utterance {u.uid}: {ref.shape} vs {res.codes[u.uid].shape}"
        assert (ref == res.codes[u.uid]).all(), f"utterance {u.uid} differs between the serving loop and a solitary run"
    alone.close()


@pytest.mark.parametrize("name", ["s1like", "projected"])
def test_producer_side_norm_statistics_against_oracle(name, monkeypatch):
    """DUALAR_TC_FUSE_NORM=2: the GEMM that produces an activation leaves the RMSNorm's sums of squares (per column and 128-row tile, from
    its epilogue), the consuming GEMM normalises its own operand with them -- no norm kernel between the two.  The fp32 order of the
    sum of squares differs from the norm kernel's, so the check is the oracle replay with the usual tolerance, and the kernel count."""
    monkeypatch.setenv("DUALAR_TC_FUSE_NORM", "2")
    cfg = variant_configs()[name]
    sd = make_state_dict(cfg, seed=0)
    m = orc.OracleModel.build(cfg, sd, device="cuda:0")
    eng = DualAREngine(cfg, sd, device=0)
    eng.batch_init(5, cfg.max_seq_len)
    n = 10
    prompts, cols, slows, fasts = run_batch(eng, cfg, REQS, n)
    launches = int(eng.batch_read("launches")[0])
    same = 0
    for sl, r in enumerate(REQS):
        same += replay_slot(cfg, m, eng, prompts[sl], cols[sl], slows[sl], fasts[sl], r[3], r[4], r[5], 100 + sl, f"{name} fused slot {sl}")
    eng.close()
    monkeypatch.setenv("DUALAR_TC_FUSE_NORM", "0")
    eng = DualAREngine(cfg, sd, device=0)
    eng.batch_init(5, cfg.max_seq_len)
    plain = int(eng.batch_read("launches")[0])
    eng.close()
    print(f"\n[{name}] {launches} kernels per step with producer-side statistics, {plain} without; {same}/{n * len(REQS)} (slot, step) pairs identical to the oracle")
    assert launches < plain - cfg.n_layer and same >= int(0.8 * n * len(REQS))


def test_tensor_core_attention_against_oracle(monkeypatch):
    """The decode attention's Q.K^T and P@V as mma.sync m16n8k16 (default; bf16 x bf16 products are exact in fp32, P split exactly into
    three bf16 terms, K / V tiles by 2-D TMA with the 128-byte swizzle) against the scalar fp32 walk (DUALAR_ATTN_MMA=0): oracle replay
    with the usual tolerance, and the two kernels' logits next to each other."""
    cfg = variant_configs()["s1like"]
    sd = make_state_dict(cfg, seed=0)
    m = orc.OracleModel.build(cfg, sd, device="cuda:0")
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("DUALAR_ATTN_MMA", flag)
        eng = DualAREngine(cfg, sd, device=0)
        eng.batch_init(5, cfg.max_seq_len)
        prompts, cols, slows, fasts = run_batch(eng, cfg, REQS, 10)
        if flag == "1":
            same = sum(replay_slot(cfg, m, eng, prompts[sl], cols[sl], slows[sl], fasts[sl], r[3], r[4], r[5], 100 + sl, f"mma slot {sl}") for sl, r in enumerate(REQS))
            assert same >= int(0.8 * 10 * len(REQS))
        outs[flag] = (cols, slows)
        eng.close()
    worst = max(float((outs["1"][1][sl][s].float() - outs["0"][1][sl][s].float()).abs().max()) for sl in range(5) for s in range(10))
    print(f"\n[tensor-core attention] max |slow logit (mma) - slow logit (scalar)| over 5 slots x 10 steps: {worst:.4f}")
    assert worst < 0.13      # one bf16 ulp at |x| <= 16 is 0.125 ... 0.0625: accumulation-order noise


def test_continuous_batching_refill_does_not_disturb_neighbours():
    cfg = variant_configs()["s1like"]
    sd = make_state_dict(cfg, seed=0)
    eng = DualAREngine(cfg, sd, device=0)
    eng.batch_init(4, cfg.max_seq_len)
    _, ref_cols, _, _ = run_batch(eng, cfg, REQS[:4], 16)
    for sl in range(4):
        eng.batch_release(sl)
    # same four requests, but slot 1 is limited to 5 tokens, collected, released and refilled with request 4 after 8 steps
    prompts = [make_prompt(cfg, r) for r in REQS]
    for sl, r in enumerate(REQS[:4]):
        eng.batch_prefill(sl, prompts[sl], 5 if sl == 1 else 16, r[3], r[4], r[5], seed=100 + sl)
    eng.batch_decode(8)
    short, fin = eng.batch_collect(1)
    assert fin and short.shape[1] == 5 and (short == ref_cols[1][:, :5]).all()
    eng.batch_release(1)
    r = REQS[4]
    eng.batch_prefill(1, prompts[4], 8, r[3], r[4], r[5], seed=777)
    eng.batch_decode(8)
    for sl in (0, 2, 3):
        cols, fin = eng.batch_collect(sl)
        assert fin and (cols == ref_cols[sl]).all(), f"slot {sl} was disturbed by the refill of slot 1"
    newcols, fin = eng.batch_collect(1)
    assert fin and newcols.shape[1] == 8
    # the refilled request equals the same request run in an otherwise empty batch
    for sl in range(4):
        eng.batch_release(sl)
    eng.batch_prefill(3, prompts[4], 8, r[3], r[4], r[5], seed=777)
    eng.batch_decode(8)
    alone, _ = eng.batch_collect(3)
    eng.close()
    assert (alone == newcols).all()


@pytest.mark.parametrize("name", list(variant_configs().keys()))
def test_tensor_core_prefill_kv_rows(name):
    """dualar_prefill (prefill_mode 0: positions [0, T-1) through the tcgen05 GEMMs) writes the KV rows the oracle's one-shot
    prefill writes, and the round-1 one-position-per-launch prefill (prefill_mode 1) agrees with it too"""
    cfg = variant_configs()[name]
    sd = make_state_dict(cfg, seed=0)
    prompt = synthetic_prompt(cfg, 6, 70, 4, seed=3)        # 80+ positions: more than one 64-row attention tile
    Tlen = prompt.size(1)
    ref = orc.OracleModel.build(cfg, sd, device="cuda:0")
    ref.setup_caches(cfg.max_seq_len)
    t = [torch.tensor(v, device="cuda:0", dtype=torch.float) for v in (0.7, 0.8, 1.1)]
    with torch.inference_mode(), sdpa_kernel(SDPBackend.MATH):
        orc.decode_one_token_ar(ref, prompt.cuda().view(1, cfg.num_codebooks + 1, -1), torch.arange(Tlen, device="cuda:0"), *t, None, stable_ties=True)
    firsts = []
    for mode in (0, 1):
        m, eng, _ = build_pair(cfg, seed=0, sd=sd)              # engine bound to m's KV tensors
        eng.set_option("prefill_mode", mode)
        eng.prefill(prompt, 4, 0.7, 1e-9, 1.0)
        cols, _ = eng.collect()
        firsts.append((int(cols[0, 0]), eng.read("slow_logits_raw").clone()))
        for l in range(cfg.n_layer):
            for which in (0, 1):
                mine, want = m.kv[l][which][0, :, :Tlen].float(), ref.kv[l][which][0, :, :Tlen].float()
                # K rows sit behind the q/k RMSNorm, which turns one bf16 ulp of the wqkv output into a few: 4 ulp, and a tight mean
                d = (mine - want).abs()
                lim = torch.maximum(torch.full_like(d, 4e-2), 4 * bf16_ulp(want))
                assert (d <= lim).all(), f"{name} mode {mode} layer {l} {'kv'[which]}: worst {d.max().item():.4f}"
                assert d.mean().item() < 2e-3, f"{name} mode {mode} layer {l} {'kv'[which]}: mean {d.mean().item():.5f}"
        eng.close()
    logits_close(firsts[0][1], firsts[1][1], cfg, f"{name}: first-token logits, tensor-core prefill vs replay prefill")
    if firsts[0][0] != firsts[1][0]:
        assert near_tie(firsts[1][1], firsts[0][0], firsts[1][0])


def test_s1_mini_batched_decode_against_oracle():
    """full-size model, 4 slots with different prompts and sampling parameters, replayed slot by slot through the oracle"""
    cfg = s1_mini_config()
    sd = make_state_dict(cfg, seed=0)
    m = orc.OracleModel.build(cfg, sd, device="cuda:0")
    eng = DualAREngine(cfg, sd, device=0)
    eng.batch_init(4, 512)
    n = 6
    reqs = [(5, 100, 1, 0.7, 0.8, 1.1), (3, 30, 2, 0.7, 1e-9, 1.0), (7, 215, 3, 1.0, 1.0, 1.5), (4, 12, 4, 0.7, 0.7, 1.5)]
    prompts, cols, slows, fasts = run_batch(eng, cfg, reqs, n)
    sem = slice(cfg.semantic_begin_id, cfg.semantic_end_id + 1)
    worst = 0.0

    def tol(o):
        nonlocal worst
        d = (o["my_slow"].float() - o["ref_slow"].float()).abs()
        worst = max(worst, d[sem].max().item())
        lim = torch.maximum(torch.full_like(d, 5e-2), 8 * bf16_ulp(o["ref_slow"]))
        assert (d <= lim).all(), f"logits beyond max(5e-2, 8 bf16 ulp): worst {d.max().item()}"
        assert d[sem].mean().item() < 1e-2

    same = 0
    for sl, r in enumerate(reqs):
        same += replay_slot(cfg, m, eng, prompts[sl], cols[sl], slows[sl], fasts[sl], r[3], r[4], r[5], 100 + sl, f"s1-mini slot {sl}",
                            slow_tol=tol, atol=7e-2, ulps=8.0)
    launches = int(eng.batch_read("launches")[0])
    eng.close()
    print(f"\n[s1-mini batched] {launches} kernels per step; semantic-logit max |ours - oracle| = {worst:.4f}; {same}/{n * 4} (slot, step) pairs identical in all rows")
    assert same >= int(0.6 * n * 4)


def test_s1_mini_serving_loop_is_deterministic_and_group_invariant():
    """BASELINE-size property test of the serving loop: 48 mixed utterances (prompts 64..200, 16..48 tokens) on the full-size model
    through 64 request slots in two concurrent groups, twice -- identical codes (no race between groups, asynchronous prefills and
    refills) -- and once more through ONE group of 32 slots: the grouping does not change a single code either.  Ids in range."""
    from fish_tts_b200 import replicas
    cfg = s1_mini_config()
    sd = make_state_dict(cfg, seed=0)
    rng = np.random.default_rng(5)
    utts = []
    for i in range(48):
        T = int(rng.integers(64, 201))
        utts.append(replicas.Utterance(uid=i, prompt=synthetic_prompt(cfg, 3, T - 8, 5, seed=2000 + i).numpy(), max_new_tokens=int(rng.integers(16, 49))))
    runs = []
    for max_batch, gs in ((64, 32), (64, 32), (32, 32)):
        eng = DualAREngine(cfg, sd, device=0)
        eng.batch_init(max_batch, 320, group_slots=gs)
        res = replicas.run_rank_batched(eng, utts, 0, 1, max_batch, poll_steps=8, sync=torch.cuda.synchronize, seed_base=9000)
        eng.close()
        assert sorted(res.uids) == list(range(48)) and res.tokens == sum(u.max_new_tokens for u in utts)
        runs.append(res.codes)
    for u in utts:
        a, b, c = (r[u.uid] for r in runs)
        assert a.shape == (cfg.num_codebooks + 1, u.max_new_tokens)
        assert (a == b).all(), f"utterance {u.uid}: two identical runs differ"
        assert (a == c).all(), f"utterance {u.uid}: two groups of 32 and one group of 32 differ"
        assert ((a[0] >= cfg.semantic_begin_id) & (a[0] <= cfg.semantic_end_id)).all() and ((a[1:] >= 0) & (a[1:] < cfg.codebook_size)).all()


@pytest.mark.parametrize("max_batch", [9, 3])
def test_long_context_slots_split_kv_paths(max_batch):
    """Contexts of many 64-position tiles: with more than 8 slots the KV range is split from 8 tiles on and merged through partials + a
    ticket (the last split owns the new position: it computes, writes and patches the K / V row); with at most 8 slots every tile is a
    split of a 4-CTA cluster merging through distributed shared memory.  Both against the oracle replay, next to short requests."""
    cfg = tiny_config(max_seq_len=2048)
    sd = make_state_dict(cfg, seed=0)
    m = orc.OracleModel.build(cfg, sd, device="cuda:0")
    eng = DualAREngine(cfg, sd, device=0)
    eng.batch_init(max_batch, 2048)
    reqs = [(5, 1200, 1, 0.7, 0.8, 1.1), (3, 30, 2, 0.7, 1e-9, 1.0), (4, 600, 4, 0.3, 0.5, 1.2)]
    n = 6
    prompts, cols, slows, fasts = run_batch(eng, cfg, reqs, n)
    same = 0
    for sl, r in enumerate(reqs):
        same += replay_slot(cfg, m, eng, prompts[sl], cols[sl], slows[sl], fasts[sl], r[3], r[4], r[5], 100 + sl, f"long slot {sl} ({max_batch} slots)")
    eng.close()
    assert same >= int(0.8 * n * len(reqs))
