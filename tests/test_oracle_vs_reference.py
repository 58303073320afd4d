"""Live pin: oracle vs the unmodified reference imported from /root/reference (skipped where it is absent,
e.g. on the GPU box)."""
import tempfile

import pytest
import torch

from fish_tts_b200 import philox
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt, weight_manifest
from helpers import variant_configs
from oracle import dualar_oracle as orc
from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not present")


@pytest.mark.parametrize("name", list(variant_configs().keys()))
def test_bit_exact_against_reference(name):
    cfg = variant_configs()[name]
    sd = make_state_dict(cfg, seed=3)
    _, inf = rh.import_reference()
    with tempfile.TemporaryDirectory() as d:
        model, _ = rh.load_reference_model(cfg, sd, d)
    assert set(model.state_dict().keys()) == {k for k, _, _ in weight_manifest(cfg)}
    prompt = synthetic_prompt(cfg, 3, 9, 5, seed=4)
    for (T, p, rp) in [(0.7, 0.8, 1.1), (0.7, 1e-9, 1.0)]:
        model._cache_setup_done = False
        model.max_seq_len = -1
        with rh.RecordingStep(model, inf, philox.oracle_noise_fn(cfg, 5)) as rec:
            y = inf.generate(model=model, prompt=prompt.clone(), max_new_tokens=24, audio_masks=None, audio_parts=None,
                             decode_one_token=rec.step, temperature=T, top_p=p, repetition_penalty=rp)
        m = orc.OracleModel.build(cfg, sd)
        tr = []
        y2 = orc.generate(m, prompt.clone(), 24, T, p, rp, noise=orc.NoiseSource(philox.oracle_noise_fn(cfg, 5)), trace=tr)
        assert torch.equal(y, y2)
        for a, b in zip(rec.steps, tr):
            assert torch.equal(a["slow_logits"], b.slow_logits)
            assert torch.equal(a["hidden"], b.hidden)
            assert all(torch.equal(x, z) for x, z in zip(a["fast_logits"], b.fast_logits))


def test_reference_rng_path_is_reproduced():
    """without explicit noise both draw from torch's global generator in the same order"""
    cfg = variant_configs()["s1like"]
    sd = make_state_dict(cfg, seed=3)
    _, inf = rh.import_reference()
    with tempfile.TemporaryDirectory() as d:
        model, decode = rh.load_reference_model(cfg, sd, d)
    prompt = synthetic_prompt(cfg, 3, 9, 5, seed=4)
    torch.manual_seed(77)
    y = inf.generate(model=model, prompt=prompt.clone(), max_new_tokens=16, audio_masks=None, audio_parts=None,
                     decode_one_token=decode, temperature=0.7, top_p=0.8, repetition_penalty=1.1)
    torch.manual_seed(77)
    y2 = orc.generate(orc.OracleModel.build(cfg, sd), prompt.clone(), 16, 0.7, 0.8, 1.1)
    assert torch.equal(y, y2)


def test_prompt_packing_equals_encode_for_inference():
    """fish_tts_b200.prompt.pack_prompt (numpy, whole arrays) == ContentSequence.encode_for_inference (inference.py:611-640, one
    .item() per code) on mixed text / VQ parts, for one request and for a ragged batch"""
    import numpy as np

    from fish_tts_b200.prompt import pack_prompt, pack_prompts
    cfg = variant_configs()["s1like"]
    sd = make_state_dict(cfg, seed=3)
    _, inf = rh.import_reference()
    with tempfile.TemporaryDirectory() as d:
        model, _ = rh.load_reference_model(cfg, sd, d)
    tok = model.tokenizer
    rng = np.random.default_rng(0)
    reqs = []
    for n_parts in (1, 2, 5, 4):
        parts = []
        for i in range(n_parts):
            if i % 2 == 0:
                parts.append(rng.integers(0, 200, size=int(rng.integers(1, 9))).astype(np.int64))
            else:
                parts.append(rng.integers(0, cfg.codebook_size, size=(cfg.num_codebooks, int(rng.integers(1, 30)))).astype(np.int64))
        reqs.append(parts)
    for parts in reqs:
        seq = inf.ContentSequence()
        for p in parts:
            seq.append(inf.TextPart(tokens=p.tolist()) if p.ndim == 1 else inf.VQPart(codes=torch.from_numpy(p)), add_end=False)
        ref, masks, audio = seq.encode_for_inference(tok, cfg.num_codebooks)
        mine = pack_prompt(parts, cfg.num_codebooks, tok.semantic_begin_id, cfg.codebook_size)
        assert mine.dtype == np.int32 and mine.shape == tuple(ref.shape) and (torch.from_numpy(mine) == ref).all()
    batch, lens = pack_prompts(reqs, cfg.num_codebooks, tok.semantic_begin_id)
    assert [b.shape[1] for b in batch] == lens.tolist() and all((b == pack_prompt(r, cfg.num_codebooks, tok.semantic_begin_id)).all() for b, r in zip(batch, reqs))
    with pytest.raises(ValueError):
        pack_prompt([np.zeros((cfg.num_codebooks + 1, 3))], cfg.num_codebooks, tok.semantic_begin_id)
