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


@pytest.mark.parametrize("dist", ["flat", "narrow", "peaks", "ties", "equal", "dominant", "tail", "split"])
def test_binned_nucleus_equals_sorted_prefix(dist):
    """The sort-free nucleus search of sample_binned() keeps exactly the prefix the definition keeps (numpy models of both;
    the CUDA implementations are compared on the GPU by tests/cuda/binned_check.cu)."""
    import numpy as np
    rng = np.random.default_rng(hash(dist) % (2 ** 32))
    for V in (1024, 4096):
        for top_p in (1e-9, 0.05, 0.3, 0.8, 0.95, 1.0):
            for s_factor in (1.0, 3.0):          # 3.0: candidates of a larger vocabulary (S covers more than the items)
                u = rng.random(V).astype(np.float32)
                i = np.arange(V)
                z = {"flat": (u - 0.5) * 6, "narrow": u * 3, "peaks": np.where(i % 97 == 0, 9 + u, u * 2), "ties": np.floor(u * 8) * 0.25,
                     "equal": np.full(V, 1.5), "dominant": np.where(i == 123, 20.0, (u - 0.5) * 4), "tail": -30 * u,
                     "split": np.where(u < 0.5, 2 + 0.001 * u, -50.0)}[dist].astype(np.float32)
                z = torch.from_numpy(z).bfloat16().float().numpy()
                e = np.exp(z - z.max()).astype(np.float32)
                S = np.float32(e.sum(dtype=np.float64)) * np.float32(s_factor)
                p = torch.from_numpy(e / S).bfloat16().float().numpy()
                w = (p.astype(np.float64) * sm.FIX).astype(np.int64)
                c_max = sm.cmax_from_top_p(top_p)
                a, b = sm.nucleus_sorted(z, w, c_max), sm.nucleus_binned(z, w, c_max)
                assert (a == b).all(), f"{dist} V={V} top_p={top_p} S x{s_factor}: {int(a.sum())} vs {int(b.sum())} kept"


def test_where_the_three_nucleus_definitions_can_differ():
    """DESIGN.md section 5.  The nucleus is {j : bf16(cum_j) <= bf16(top_p)} over the sorted bf16 probabilities; three
    definitions of cum_j are in play:
      (R) the reference: torch.cumsum on the bf16 tensor (CPU: an fp32 running sum; CUDA: a parallel scan in unspecified order),
      (O) the oracle with stable_ties: an fp32 running sum in sorted order, rounded to bf16 per element (= R on the CPU),
      (K) the kernels: the EXACT sum (2^-44 fixed point), compared against the bf16 rounding boundary of top_p.
    bf16(cum_j) only changes at a bf16 rounding boundary, and the cut is decided at ONE boundary: M = the midpoint between
    bf16(top_p) and its successor.  K and O/R can therefore only disagree when the exact cumulative sum at the cut passes M
    within the rounding error of an fp32 running sum of n terms <= 1 (n * 2^-24).  Checked here on random vectors: whenever
    the nucleus sizes differ, the exact sum at the boundary index lies within that error of M; otherwise they are identical --
    and they are identical in the overwhelming majority of cases."""
    import numpy as np
    rng = np.random.default_rng(123)
    differ, total = 0, 0
    for V in (1024, 4096):
        for scale in (0.3, 0.64, 2.5):
            for top_p in (0.3, 0.7, 0.8, 0.95):
                for _ in range(12):
                    z = torch.from_numpy((rng.standard_normal(V) * scale).astype(np.float32)).bfloat16()
                    zs, _ = torch.sort(z, descending=True, stable=True)
                    p = torch.softmax(zs, dim=-1)                                         # bf16 probabilities, the reference's own op
                    tp = torch.tensor(top_p).bfloat16()
                    n_ref = int((torch.cumsum(p, dim=-1) <= tp).sum())                    # (R) on the CPU
                    n_orc = int((torch.cumsum(p.float(), dim=-1).to(torch.bfloat16) <= tp).sum())      # (O)
                    w = (p.float().double().numpy() * sm.FIX).astype(np.int64)
                    cum = np.cumsum(w)
                    c_max = sm.cmax_from_top_p(top_p)
                    n_k = int((cum <= c_max).sum())                                       # (K)
                    total += 1
                    assert n_ref == n_orc, "stable_ties keeps the reference's CPU nucleus (cumsum order is the sorted order on both)"
                    if n_k != n_orc:
                        differ += 1
                        j = min(n_k, n_orc)                                               # first index on which they disagree
                        gap = abs(int(cum[min(j, V - 1)]) - c_max) / sm.FIX
                        assert abs(n_k - n_orc) <= 2 and gap <= V * 2.0 ** -23 + float(p[min(j, V - 1)]), \
                            f"nucleus sizes {n_k} vs {n_orc} differ away from the rounding boundary (gap {gap:.3e})"
    assert differ <= 0.05 * total, f"{differ}/{total} vectors with a boundary disagreement"


def test_three_way_bf16_split_of_fp32_is_exact():
    """The tensor-core attention feeds P (fp32 probabilities) to bf16 MMAs as hi + mid + lo, each the round-to-nearest bf16 of what is
    left (batch.cuh, b_attn_body<B, true>): 24 significant bits = 3 x 8, so the three terms add up to P exactly and the products with
    bf16 V values are the fp32 products.  Checked on random values, on the edges of binades and on tiny / subnormal-adjacent inputs."""
    import torch
    g = torch.Generator().manual_seed(0)
    x = torch.cat([torch.rand(200_000, generator=g), torch.rand(50_000, generator=g) * 1e-6,
                   torch.tensor([1.0, 0.5, 0.99999994, 0.50000006, 2.0 ** -20, 3.0 * 2.0 ** -30, 1.0 - 2.0 ** -9, 1.0 - 2.0 ** -17]),
                   torch.exp(-torch.rand(100_000, generator=g) * 30.0)]).to(torch.float32)

    def rbf(t):
        return t.to(torch.bfloat16).to(torch.float32)
    hi = rbf(x); r1 = x - hi
    mid = rbf(r1); r2 = r1 - mid
    lo = rbf(r2)
    assert torch.equal(r2, lo), "the third term must capture the remainder exactly"
    assert torch.equal((hi.double() + mid.double() + lo.double()).to(torch.float32), x)
    assert torch.equal(hi.double() + mid.double() + lo.double(), x.double())
