"""Shared helpers of the parity tests (oracle side = test infrastructure)."""
from __future__ import annotations

import torch

from fish_tts_b200.config import DualARConfig, tiny_config


def bf16_ulp(x: torch.Tensor) -> torch.Tensor:
    """one bf16 ulp at |x| (float tensor)"""
    ax = x.float().abs().clamp_min(2.0 ** -126)
    return torch.exp2(torch.floor(torch.log2(ax)) - 7)


def logits_close(mine: torch.Tensor, ref: torch.Tensor, cfg: DualARConfig | None = None, where: str = "",
                 atol: float = 2e-2, ulps: float = 2.0):
    """The north-star tolerance: per-step logits within 2e-2 max-abs (bf16).  bf16 cannot represent
    a difference below one ulp, which exceeds 2e-2 once |x| >= 4, so the bound used is
    max(2e-2, 2 ulp(ref)); over the semantic range (the only ids that can be sampled from the
    conditioned checkpoints, |logit| < 4) this is the literal 2e-2."""
    mine, ref = mine.float().cpu(), ref.float().cpu()
    diff = (mine - ref).abs()
    tol = torch.maximum(torch.full_like(diff, atol), ulps * bf16_ulp(ref))
    bad = diff > tol
    assert not bad.any(), f"{where}: {int(bad.sum())} logits off; worst {diff.max().item():.4f} at {int(diff.argmax())} (ref {ref.flatten()[diff.argmax()].item():.4f})"
    if cfg is not None and mine.numel() == cfg.vocab_size:
        sem = slice(cfg.semantic_begin_id, cfg.semantic_end_id + 1)
        assert diff[sem].max().item() <= atol or ref[sem].abs().max().item() >= 4, f"{where}: semantic logits off by {diff[sem].max().item()}"
    return diff.max().item(), float((diff > 0).float().mean())


def near_tie(ref_logits: torch.Tensor, tok_mine: int, tok_ref: int, ulps: float = 2.0, atol: float = 0.0) -> bool:
    """greedy disagreement is explainable iff the reference itself scores the two tokens within twice the per-logit
    tolerance max(atol, ulps * bf16 ulp) -- each of the two logits may be off by that much"""
    a, b = ref_logits[tok_mine].float(), ref_logits[tok_ref].float()
    tol = torch.maximum(torch.tensor(atol), ulps * bf16_ulp(torch.maximum(a.abs(), b.abs())))
    return bool((a - b).abs() <= 2 * tol)


def variant_configs():
    """tiny shapes that together switch every config branch of the path"""
    return {
        "s1like": tiny_config(),                                           # tied head, qk-norm, GQA 2, scaled codebook emb
        "v15like": tiny_config(tie_word_embeddings=False, attention_qk_norm=False, scale_codebook_embeddings=False,
                               n_head=8, n_local_heads=1, head_dim=32, fast_n_head=8, fast_n_local_heads=1,
                               fast_head_dim=32, num_codebooks=8),
        "biased": tiny_config(attention_qkv_bias=True, attention_o_bias=True, fast_attention_qk_norm=True,
                              n_layer=2, head_dim=64, n_head=4, n_local_heads=4),
        "projected": tiny_config(fast_dim=512, fast_n_head=8, fast_n_local_heads=4, fast_head_dim=64,             # fast_dim != dim:
                                 fast_intermediate_size=768, n_layer=2),                                          # fast_project_in (llama.py:510-513)
    }
