"""Pin the oracle (oracle/dualar_oracle.py) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Same torch build and CPU => bit-identical; on a different host CPU the
bf16 GEMM summation order may differ, so the fallback bound is the north-star tolerance."""
from pathlib import Path

import pytest
import torch

from fish_tts_b200 import philox
from fish_tts_b200.config import s1_mini_config
from fish_tts_b200.synthetic import make_state_dict
from helpers import logits_close, variant_configs
from oracle import dualar_oracle as orc

GOLD = Path(__file__).resolve().parent / "golden"


def run_oracle(cfg, sd, g, stable=False):
    m = orc.OracleModel.build(cfg, sd)
    tr = []
    seq = orc.generate(m, g["prompt"].clone(), g["seq"].size(1) - g["prompt"].size(1) if "eos" not in str(g.get("tag", "")) else 40,
                       g["temperature"], g["top_p"], g["repetition_penalty"],
                       noise=orc.NoiseSource(philox.oracle_noise_fn(cfg, g["noise_seed"])), stable_ties=stable, trace=tr)
    return seq, tr


@pytest.mark.parametrize("name", list(variant_configs().keys()))
@pytest.mark.parametrize("mode", ["sampled", "greedy", "hot"])
def test_tiny_matches_reference(name, mode):
    cfg = variant_configs()[name]
    g = torch.load(GOLD / f"tiny_{name}_{mode}.pt")
    sd = make_state_dict(cfg, seed=0)
    seq, tr = run_oracle(cfg, sd, g)
    exact = torch.equal(seq, g["seq"])
    for i, t in enumerate(tr):
        logits_close(t.slow_logits, g["slow_logits"][i], cfg, f"{name}/{mode} step {i} slow")
        logits_close(torch.stack(t.fast_logits), g["fast_logits"][i], None, f"{name}/{mode} step {i} fast")
        if not torch.equal(t.tokens.cpu(), g["tokens"][i]):
            break   # trajectories may legitimately part ways after a near-tie on another CPU
    if not exact:
        pytest.xfail("oracle == reference bit-for-bit only on the torch build / CPU ISA the fixtures were made on")
    assert all(torch.equal(t.slow_logits, g["slow_logits"][i]) for i, t in enumerate(tr))


@pytest.mark.parametrize("name", list(variant_configs().keys()))
def test_tiny_early_stop(name):
    """<|im_end|> ends the loop after the column is recorded (inference.py:208-211)."""
    cfg = variant_configs()[name]
    g = torch.load(GOLD / f"tiny_{name}_eos.pt")
    g["tag"] = "eos"
    sd = make_state_dict(cfg, seed=0, eos_reachable=True)
    seq, _ = run_oracle(cfg, sd, g)
    assert seq.shape == g["seq"].shape and int(seq[0, -1]) == cfg.im_end_id
    assert torch.equal(seq, g["seq"])


def test_stable_ties_changes_nothing_when_no_boundary_tie():
    """the `stable_ties` knob (index-ordered ties, fp32 running sum) reproduces the reference on these cases"""
    cfg = variant_configs()["s1like"]
    g = torch.load(GOLD / "tiny_s1like_sampled.pt")
    sd = make_state_dict(cfg, seed=0)
    seq, _ = run_oracle(cfg, sd, g, stable=True)
    same = (seq == g["seq"]).all(0).float().mean().item()
    assert same > 0.9


@pytest.mark.parametrize("mode", ["sampled", "greedy"])
def test_s1_mini_matches_reference(mode):
    cfg = s1_mini_config()
    g = torch.load(GOLD / f"s1mini_{mode}.pt")
    sd = make_state_dict(cfg, seed=0)
    seq, tr = run_oracle(cfg, sd, g)
    for i, t in enumerate(tr):
        logits_close(t.slow_logits[::16], g["slow_strided"][i], None, f"s1mini/{mode} step {i} strided")
        logits_close(t.slow_logits[g["slow_top_idx"][i].long()], g["slow_top_val"][i], None, f"s1mini/{mode} step {i} top")
        logits_close(torch.stack(t.fast_logits), g["fast_logits"][i], None, f"s1mini/{mode} step {i} fast")
        if not torch.equal(t.tokens.cpu(), g["tokens"][i]):
            break
    if not torch.equal(seq, g["seq"]):
        pytest.xfail("bit-exactness holds on the torch build / CPU ISA the fixtures were made on")
