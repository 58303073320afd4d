"""The CUDA sampler's definition (sort-free, fixed-point nucleus, index-ordered ties) restated in numpy
(oracle/sampler_model.py) agrees with the reference's sampler (inference.py:30-80) on the CPU."""
import pytest
import torch

from fish_tts_b200 import philox
from oracle import dualar_oracle as orc
from oracle import sampler_model as sm

CASES = [(0.7, 0.8, 1.1), (0.7, 0.7, 1.5), (1.0, 1.0, 1.0), (0.7, 1e-9, 1.0), (0.3, 0.5, 1.2), (1.9, 0.99, 0.5)]


@pytest.mark.parametrize("V", [640, 1024, 4096, 20000])
@pytest.mark.parametrize("scale", [0.64, 2.5, 6.0])
def test_definition_matches_reference_sampler(V, scale):
    g = torch.Generator().manual_seed(V + int(scale * 10))
    for i, (T, p, rp) in enumerate(CASES):
        logits = (torch.randn(V, generator=g) * scale).bfloat16()
        noise = philox.exp1_noise(9, i, 0, V)
        prev = torch.randint(0, V, (11,), generator=g)
        tok, probs = orc.sample(logits.clone().view(1, 1, -1), torch.tensor(T), torch.tensor(p), torch.tensor(rp), prev,
                                noise=orc.NoiseSource(lambda c, n: noise), stable_ties=True)
        tok2, nk = sm.sample(logits, T, p, rp, prev.numpy(), noise)
        assert abs(int((probs > 0).sum()) - nk) <= 1, "nucleus size"
        assert int(tok) == tok2


def test_all_equal_logits_and_extremes():
    V = 640
    noise = philox.exp1_noise(1, 0, 0, V)
    for logits in (torch.zeros(V), torch.full((V,), -3.0), torch.cat([torch.full((1,), 30.0), torch.zeros(V - 1)])):
        logits = logits.bfloat16()
        for (T, p, rp) in CASES:
            tok, _ = orc.sample(logits.clone().view(1, 1, -1), torch.tensor(T), torch.tensor(p), torch.tensor(rp), None,
                                noise=orc.NoiseSource(lambda c, n: noise), stable_ties=True)
            tok2, _ = sm.sample(logits, T, p, rp, None, noise)
            assert int(tok) == tok2


def test_philox_known_answers():
    c = philox.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(x[0]) for x in c] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    c = philox.philox4x32_10([0xFFFFFFFF] , [0xFFFFFFFF], [0xFFFFFFFF], [0xFFFFFFFF], 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(x[0]) for x in c] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    q = philox.exp1_noise(7, 0, 0, 200000).float()
    assert abs(q.mean().item() - 1.0) < 0.02 and q.min().item() > 0
