"""Host-side logic that needs no GPU: config mirror, weight manifest, partitioner, the 2-rank (gloo) replica path."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from fish_tts_b200 import replicas
from fish_tts_b200.config import DualARConfig, fish_speech_1_5_config, s1_mini_config, tiny_config
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt, weight_manifest

ROOT = Path(__file__).resolve().parent.parent


def test_s1_mini_byte_budget_matches_survey():
    """SURVEY.md section 8d / BASELINE.md section 3: the numbers the roofline is computed from"""
    w = s1_mini_config().weight_bytes()
    assert abs(w["slow_layers"] / 1e6 - 880.93) < 0.1
    assert abs(w["lm_head"] / 1e6 - 319.03) < 0.1
    assert abs(w["fast_layers"] / 1e6 - 100.68) < 0.1
    assert w["kv_per_pos"] == 114688
    assert abs(s1_mini_config().algorithmic_bytes_per_token(732) / 1e9 - 1.393) < 0.005
    n = sum(int(np.prod(s)) for _, s, _ in weight_manifest(s1_mini_config()))
    assert n == 700_654_592                       # SURVEY.md probe table
    n15 = sum(int(np.prod(s)) for _, s, _ in weight_manifest(fish_speech_1_5_config()))
    assert n15 == 637_921_280


def test_config_defaults_follow_reference_rules(tmp_path):
    c = DualARConfig(dim=1024, n_head=16, head_dim=None, intermediate_size=None, n_local_heads=-1)
    assert c.head_dim == 64 and c.n_local_heads == 16 and c.intermediate_size == 2816 and c.fast_dim == 1024
    (tmp_path / "config.json").write_text('{"model_type": "dual_ar", "dim": 512, "n_head": 8, "unknown_key": 1}')
    assert DualARConfig.from_json(tmp_path, im_end_id=4).dim == 512
    (tmp_path / "config.json").write_text('{"model_type": "naive"}')
    with pytest.raises(ValueError):
        DualARConfig.from_json(tmp_path)


def test_synthetic_checkpoint_is_deterministic_and_conditioned():
    cfg = tiny_config()
    a, b = make_state_dict(cfg, seed=7), make_state_dict(cfg, seed=7)
    assert all(torch.equal(a[k], b[k]) for k in a)
    head = a["embeddings.weight"]
    assert (head[: cfg.semantic_begin_id, 0] < 0).all() and (head[cfg.semantic_begin_id: cfg.semantic_end_id + 1, 0] == 0).all()
    p = synthetic_prompt(cfg, 5, 12, 4)
    assert p.shape == (cfg.num_codebooks + 1, 21) and p.dtype == torch.int32
    vq = p[:, 5:17]
    assert ((vq[0] >= cfg.semantic_begin_id) & (vq[0] <= cfg.semantic_end_id)).all() and (vq[1] == vq[0] - cfg.semantic_begin_id).all()


def test_partition_is_exact_and_balanced():
    rng = np.random.default_rng(0)
    costs = rng.integers(128, 1024, size=4096).astype(float)
    for world in (1, 2, 4, 8):
        parts = replicas.partition_longest_first(costs, world)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(4096))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) / (sum(loads) / world) < 1.01
    assert replicas.partition_longest_first([], 4) == [[], [], [], []]
    assert replicas.partition_longest_first([5.0], 3) == [[0], [], []]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = tiny_config()
    utts = replicas.synthetic_utterances(cfg, 24, seed=3, prompt_range=(16, 32), target_range=(4, 12))

    def fake_generate(u):       # stands in for engine.generate on a CPU-only box: deterministic in (uid, length)
        return np.full((cfg.num_codebooks + 1, u.max_new_tokens), u.uid, dtype=np.int32)

    dist.barrier()
    res = replicas.run_rank(fake_generate, utts, rank, world)
    agg = replicas.aggregate(res, dist)
    q.put((rank, res.uids, res.tokens, agg))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replicas_gloo():
    """world_size 2 over gloo: disjoint shares, every utterance served once, totals reduced across ranks"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    uids0, uids1 = set(got[0][1]), set(got[1][1])
    assert uids0.isdisjoint(uids1) and uids0 | uids1 == set(range(24))
    utts = replicas.synthetic_utterances(tiny_config(), 24, seed=3, prompt_range=(16, 32), target_range=(4, 12))
    assert got[0][3]["tokens"] == got[1][3]["tokens"] == sum(u.max_new_tokens for u in utts) == got[0][2] + got[1][2]
    assert got[0][3]["world"] == 2 and got[0][3]["seconds"] == got[1][3]["seconds"]


def test_inference_mirror_signatures_match_reference():
    """same names and argument lists as fish_tts/models/inference.py (checked live when the reference is present)"""
    import inspect
    from fish_tts_b200 import inference as mine
    for name in ("init_model", "decode_one_token_ar", "decode_n_tokens", "generate", "generate_streaming", "install"):
        assert callable(getattr(mine, name))
    sig = inspect.signature(mine.decode_one_token_ar)
    assert list(sig.parameters)[:9] == ["model", "x", "input_pos", "temperature", "top_p", "repetition_penalty", "audio_masks",
                                        "audio_parts", "previous_tokens"]
    from oracle import ref_harness as rh
    if rh.reference_available():
        _, ref = rh.import_reference()
        for name in ("init_model", "decode_one_token_ar", "generate", "generate_streaming"):
            rp, mp_ = inspect.signature(getattr(ref, name)).parameters, inspect.signature(getattr(mine, name)).parameters
            assert [p for p in rp if p in mp_] == [p for p in rp], f"{name}: reference parameters {list(rp)} vs {list(mp_)}"


def test_install_swaps_the_reference_module_attributes():
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("/root/reference not present")
    from fish_tts_b200 import inference as mine
    _, ref = rh.import_reference()
    orig = ref.generate
    mine.install(ref)
    try:
        assert ref.init_model is not mine.init_model and ref.init_model.__name__ == "init_model_b200"
        assert ref.decode_one_token_ar is mine.decode_one_token_ar and ref.generate is not orig
        if not torch.cuda.is_available():
            import tempfile
            cfg = tiny_config()
            d = rh.fabricate_model_dir(cfg, make_state_dict(cfg, seed=0), tempfile.mkdtemp())
            with pytest.raises(RuntimeError):      # no GPU -> the engine refuses; it must not fall back to the reference step
                ref.init_model(str(d), "cpu", torch.bfloat16, compile=True)
    finally:
        mine.uninstall(ref)
    assert ref.generate is orig


def test_token_ids_from_model_dir(tmp_path):
    from fish_tts_b200.inference import TokenIds
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("fabricating a tokenizer file uses the reference's token list")
    for cfg in (tiny_config(), s1_mini_config()):
        d = tmp_path / str(cfg.vocab_size)
        d.mkdir()
        rh.fabricate_model_dir(cfg, {}, d)
        ids = TokenIds.from_model_dir(d)
        assert (ids.semantic_begin_id, ids.semantic_end_id, ids.im_end_id) == (cfg.semantic_begin_id, cfg.semantic_end_id, cfg.im_end_id)


def test_pack_prompt_layout():
    """row 0 = token ids (VQ positions: code 0 + semantic_begin_id), rows 1.. = codes at VQ positions, 0 elsewhere (inference.py:611-640)"""
    import numpy as np

    from fish_tts_b200.prompt import pack_prompt, pack_prompts
    C, sb = 4, 1000
    text = np.array([5, 6, 7])
    codes = np.arange(C * 2).reshape(C, 2)
    out = pack_prompt([text, codes, [9]], C, sb)
    assert out.dtype == np.int32 and out.shape == (C + 1, 6)
    assert out[0].tolist() == [5, 6, 7, sb + 0, sb + 1, 9]
    assert (out[1:, :3] == 0).all() and (out[1:, 3:5] == codes).all() and (out[1:, 5] == 0).all()
    batch, lens = pack_prompts([[text], [codes, text]], C, sb)
    assert lens.tolist() == [3, 5] and batch[1].shape == (C + 1, 5)
    assert pack_prompt([], C, sb).shape == (C + 1, 0)
    import pytest
    with pytest.raises(ValueError):
        pack_prompt([np.zeros((C, 2)) + 99], C, sb, codebook_size=16)


class _FakeSlots:
    """host-side stand-in for the request-slot API: a request produces `max_new` columns, one per step it is switched on for"""
    def __init__(self, n, groups):
        import numpy as np
        self.np, self.n, self.groups = np, n, groups
        self.left = [0] * n; self.total = [0] * n; self.open = [False] * n; self.seen_done = [False] * n
        self.pending = {}            # slot -> (max_new, uid): prefilled, joins at the next decode call (asynchronous prefill)
        self.log = []
    def batch_prefill(self, slot, prompt, max_new, *a, seed=0):
        assert not self.open[slot], "prefill into an occupied slot"
        self.open[slot] = True; self.seen_done[slot] = False; self.pending[slot] = (max_new, seed)
        self.left[slot] = -1
    def batch_decode(self, n):
        for sl, (mx, seed) in list(self.pending.items()):
            self.left[sl] = mx; self.total[sl] = mx; del self.pending[sl]
        for sl in range(self.n):
            if self.open[sl] and self.left[sl] > 0:
                self.left[sl] = max(0, self.left[sl] - n)
        self.log.append("decode")
    def batch_read(self, name):
        import torch
        if name == "groups":
            return torch.tensor([self.groups], dtype=torch.int32)
        assert name == "done"
        d = [int(self.open[sl] and self.left[sl] == 0) for sl in range(self.n)]
        for sl in range(self.n):
            self.seen_done[sl] = self.seen_done[sl] or bool(d[sl])
        self.log.append("read")
        return torch.tensor(d, dtype=torch.int32)
    def batch_collect(self, slot):
        assert self.seen_done[slot], "collected before the host saw the request finished"
        self.log.append("collect")
        return self.np.zeros((3, self.total[slot]), dtype=self.np.int32), True
    def batch_release(self, slot):
        self.open[slot] = False; self.log.append("release")


def test_serving_loop_host_logic_with_fake_slots():
    """run_rank_batched: every utterance exactly once and complete, never more requests than slots, and in pipelined mode the next
    burst is enqueued BEFORE the finished requests are collected"""
    import numpy as np

    from fish_tts_b200 import replicas
    rng = np.random.default_rng(0)
    utts = [replicas.Utterance(uid=i, prompt=np.zeros((3, int(rng.integers(4, 40))), dtype=np.int32), max_new_tokens=int(rng.integers(1, 50))) for i in range(37)]
    for groups in (1, 3):
        eng = _FakeSlots(5, groups)
        res = replicas.run_rank_batched(eng, utts, 0, 1, 5, poll_steps=4)
        assert sorted(res.uids) == list(range(37)) and res.tokens == sum(u.max_new_tokens for u in utts)
        assert all(res.codes[u.uid].shape[1] == u.max_new_tokens for u in utts)
        log = eng.log
        first_collect = log.index("collect")
        if groups > 1:
            assert log[first_collect - 1] == "decode" and log[first_collect - 2] == "read", "pipelined: read, decode, then collect"
        else:
            assert log[first_collect - 1] == "read"
    # two ranks split the work without overlap
    a = replicas.run_rank_batched(_FakeSlots(4, 2), utts, 0, 2, 4)
    b = replicas.run_rank_batched(_FakeSlots(4, 2), utts, 1, 2, 4)
    assert sorted(a.uids + b.uids) == list(range(37))
