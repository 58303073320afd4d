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
