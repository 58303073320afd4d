"""Model-shape description for the dual-AR decode path.

Mirrors the reference's ``BaseModelArgs`` / ``DualARModelArgs`` dataclasses
(reference: fish_tts/models/llama.py:31-123) field for field, so a reference
``config.json`` loads unchanged, and adds the three tokenizer-derived integer
ids that enter the hot path (reference: llama.py:418-420, inference.py:123,182):
``semantic_begin_id``, ``semantic_end_id`` and ``im_end_id``.

Nothing here touches the GPU.
"""

from __future__ import annotations

import dataclasses
import json
from dataclasses import dataclass
from pathlib import Path


def find_multiple(n: int, k: int) -> int:
    return n if n % k == 0 else n + k - (n % k)


@dataclass
class DualARConfig:
    # --- BaseModelArgs (llama.py:31-62) ---
    model_type: str = "dual_ar"
    vocab_size: int = 32000
    n_layer: int = 32
    n_head: int = 32
    dim: int = 4096
    intermediate_size: int | None = None
    n_local_heads: int = -1
    head_dim: int | None = 64
    rope_base: float = 10000
    norm_eps: float = 1e-5
    max_seq_len: int = 2048
    dropout: float = 0.0
    tie_word_embeddings: bool = True
    attention_qkv_bias: bool = False
    attention_o_bias: bool = False
    attention_qk_norm: bool = False
    codebook_size: int = 160
    num_codebooks: int = 4
    use_gradient_checkpointing: bool = True
    initializer_range: float = 0.02
    is_reward_model: bool = False
    scale_codebook_embeddings: bool = False
    # --- DualARModelArgs (llama.py:89-123) ---
    n_fast_layer: int = 4
    fast_dim: int | None = None
    fast_n_head: int | None = None
    fast_n_local_heads: int | None = None
    fast_head_dim: int | None = None
    fast_intermediate_size: int | None = None
    fast_attention_qkv_bias: bool | None = None
    fast_attention_qk_norm: bool | None = None
    fast_attention_o_bias: bool | None = None
    # --- tokenizer-derived ids (tokenizer.py:84-101) ---
    semantic_begin_id: int = -1
    semantic_end_id: int = -1
    im_end_id: int = -1

    def __post_init__(self):
        # same defaulting rules as llama.py:64-72 and llama.py:102-123
        if self.n_local_heads == -1:
            self.n_local_heads = self.n_head
        if self.intermediate_size is None:
            self.intermediate_size = find_multiple(int(2 * 4 * self.dim / 3), 256)
        if self.head_dim is None:
            self.head_dim = self.dim // self.n_head
        self.fast_dim = self.fast_dim or self.dim
        self.fast_n_head = self.fast_n_head or self.n_head
        self.fast_n_local_heads = self.fast_n_local_heads or self.n_local_heads
        self.fast_head_dim = self.fast_head_dim or self.head_dim
        self.fast_intermediate_size = self.fast_intermediate_size or self.intermediate_size
        if self.fast_attention_qkv_bias is None:
            self.fast_attention_qkv_bias = self.attention_qkv_bias
        if self.fast_attention_qk_norm is None:
            self.fast_attention_qk_norm = self.attention_qk_norm
        if self.fast_attention_o_bias is None:
            self.fast_attention_o_bias = self.attention_o_bias

    # ------------------------------------------------------------------
    @property
    def codebook_dim(self) -> int:
        """Rows of one decode-step token column: semantic id + codebooks."""
        return self.num_codebooks + 1

    def reference_json(self) -> dict:
        """The dict a reference ``config.json`` holds (no tokenizer ids)."""
        d = dataclasses.asdict(self)
        for k in ("semantic_begin_id", "semantic_end_id", "im_end_id"):
            d.pop(k)
        return d

    @staticmethod
    def from_json(path: str | Path, **token_ids) -> "DualARConfig":
        path = Path(path)
        if path.is_dir():
            path = path / "config.json"
        with open(path, "r", encoding="utf-8") as f:
            data = json.load(f)
        if data.get("model_type") != "dual_ar":
            raise ValueError(f"Unknown model type: {data.get('model_type')}")
        known = {f.name for f in dataclasses.fields(DualARConfig)}
        data = {k: v for k, v in data.items() if k in known}
        data.update(token_ids)
        return DualARConfig(**data)

    # weight bytes streamed per decode step (bf16), SURVEY.md section 8(d)
    def weight_bytes(self) -> dict:
        def layer(dim, n_head, n_kv, hd, inter, qk_norm):
            return 2 * ((n_head + 2 * n_kv) * hd * dim + dim * n_head * hd + 3 * dim * inter + 2 * dim + (2 * hd if qk_norm else 0))

        slow = self.n_layer * layer(self.dim, self.n_head, self.n_local_heads, self.head_dim,
                                    self.intermediate_size, self.attention_qk_norm)
        head = 2 * self.vocab_size * self.dim
        fast = self.n_fast_layer * layer(self.fast_dim, self.fast_n_head, self.fast_n_local_heads,
                                         self.fast_head_dim, self.fast_intermediate_size, self.fast_attention_qk_norm)
        fast_head = 2 * self.codebook_size * self.fast_dim
        if self.fast_dim != self.dim:      # fast_project_in (llama.py:510-513)
            fast += 2 * (self.fast_dim * self.dim + self.fast_dim)
        kv_per_pos = self.n_layer * 2 * self.n_local_heads * self.head_dim * 2
        return {"slow_layers": slow, "lm_head": head, "fast_layers": fast,
                "fast_head": fast_head, "kv_per_pos": kv_per_pos,
                "unique_weights": slow + head + fast + fast_head}

    def algorithmic_bytes_per_token(self, context_len: float) -> float:
        """unique weights once + KV read over the context + one KV write."""
        w = self.weight_bytes()
        return w["unique_weights"] + w["kv_per_pos"] * (context_len + 1)


# -- the two shapes BASELINE.json names (SURVEY.md section 8 preamble) --------------------

# Real tokenizer layout: 151,643 BPE ranks then ALL_SPECIAL_TOKENS in order
# (tokenizer.py:52-69): <|im_end|> is special #4, semantic tokens start at special #15.
_S1_RANKS = 151_643


def s1_mini_config() -> DualARConfig:
    return DualARConfig(
        vocab_size=155776, n_layer=28, n_head=16, dim=1024, intermediate_size=3072,
        n_local_heads=8, head_dim=128, rope_base=1e6, norm_eps=1e-6, max_seq_len=8192,
        tie_word_embeddings=True, attention_qk_norm=True, codebook_size=4096, num_codebooks=10,
        scale_codebook_embeddings=True, n_fast_layer=4, fast_dim=1024, fast_n_head=16,
        fast_n_local_heads=8, fast_head_dim=64, fast_intermediate_size=3072,
        fast_attention_qk_norm=False,
        semantic_begin_id=_S1_RANKS + 15, semantic_end_id=_S1_RANKS + 15 + 4095,
        im_end_id=_S1_RANKS + 4,
    )


def fish_speech_1_5_config(n_ranks: int = 100_000) -> DualARConfig:
    # 1,024 semantic tokens; ids re-assigned by enumeration order (tokenizer.py:84-87)
    return DualARConfig(
        vocab_size=102048, n_layer=24, n_head=16, dim=1024, intermediate_size=4096,
        n_local_heads=2, head_dim=64, rope_base=1e6, norm_eps=1e-6, max_seq_len=8192,
        tie_word_embeddings=False, attention_qk_norm=False, codebook_size=1024, num_codebooks=8,
        scale_codebook_embeddings=False, n_fast_layer=4,
        semantic_begin_id=n_ranks + 15, semantic_end_id=n_ranks + 15 + 1023,
        im_end_id=n_ranks + 4,
    )


def tiny_config(**over) -> DualARConfig:
    """A small shape with every s1-mini feature switched on; CPU oracle runs it in milliseconds."""
    n_ranks = 300
    kw = dict(
        vocab_size=640, n_layer=3, n_head=4, dim=256, intermediate_size=512, n_local_heads=2,
        head_dim=64, rope_base=1e6, norm_eps=1e-6, max_seq_len=256, tie_word_embeddings=True,
        attention_qk_norm=True, codebook_size=256, num_codebooks=4, scale_codebook_embeddings=True,
        n_fast_layer=2, fast_dim=256, fast_n_head=4, fast_n_local_heads=2, fast_head_dim=64,
        fast_intermediate_size=512, fast_attention_qk_norm=False,
        semantic_begin_id=n_ranks + 15, semantic_end_id=n_ranks + 15 + 255, im_end_id=n_ranks + 4,
    )
    kw.update(over)
    return DualARConfig(**kw)
