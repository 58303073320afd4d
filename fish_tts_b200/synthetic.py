"""Seeded random-init checkpoints and synthetic prompts for the dual-AR decode path.

There is no network, so every test, the bench and the reference harness run on
random-init weights of the named architecture (BASELINE.json ``configs``).  The
state dict uses the reference's own key names (the keys of
``DualARTransformer.state_dict()``, fish_tts/models/llama.py:349-359, 511-535), so the
same tensors feed the unchanged reference (``load_state_dict``, llama.py:498), the
oracle and the CUDA engine.

Conditioning (SURVEY.md section 8c asks for one, otherwise the step degenerates or crashes):
random-init logits are flat over the whole vocabulary, but only ids inside
``[semantic_begin_id, semantic_end_id]`` keep the codebook branch alive
(llama.py:418-429) and ids above it index past ``fast_embeddings``
(inference.py:123-125).  We therefore
  * make hidden channel 0 a positive "carrier": a block of layer-0 FFN units gets
    ``w3 == w1`` (so ``silu(a)*a >= 0``) and feeds channel 0 through ``w2``;
  * give every non-semantic row of the LM-head matrix (the tied embedding table, or
    ``output.weight``) a fixed negative entry in channel 0 and shrink its other channels.
Non-semantic logits (``<|im_end|>`` and the padded rows included) then sit roughly 5-9
below the semantic ones, so sampling stays inside the semantic range and EOS is
unreachable, while semantic logits stay |x| < 4 where one bf16 ulp is <= 2^-6.
"""

from __future__ import annotations

import torch

from .config import DualARConfig

CARRIER_UNITS = 256      # layer-0 FFN units tied w3==w1
CARRIER_GAIN = 0.25      # w2[0, unit]
NONSEM_BIAS = -2.25      # head row entry in channel 0 for non-semantic ids
NONSEM_SCALE = 0.25      # other channels of non-semantic head rows are shrunk by this


def _layer_shapes(prefix: str, dim: int, n_head: int, n_kv: int, hd: int, inter: int,
                  qkv_bias: bool, o_bias: bool, qk_norm: bool):
    """(name, shape, kind) in the reference's registration order (llama.py:196-220, 183-187, 315-320)."""
    out = [(f"{prefix}.attention.wqkv.weight", ((n_head + 2 * n_kv) * hd, dim), "w")]
    if qkv_bias:
        out.append((f"{prefix}.attention.wqkv.bias", ((n_head + 2 * n_kv) * hd,), "b"))
    out.append((f"{prefix}.attention.wo.weight", (dim, n_head * hd), "w"))
    if o_bias:
        out.append((f"{prefix}.attention.wo.bias", (dim,), "b"))
    if qk_norm:
        out.append((f"{prefix}.attention.q_norm.weight", (hd,), "n"))
        out.append((f"{prefix}.attention.k_norm.weight", (hd,), "n"))
    out += [
        (f"{prefix}.feed_forward.w1.weight", (inter, dim), "w"),
        (f"{prefix}.feed_forward.w3.weight", (inter, dim), "w"),
        (f"{prefix}.feed_forward.w2.weight", (dim, inter), "w"),
        (f"{prefix}.ffn_norm.weight", (dim,), "n"),
        (f"{prefix}.attention_norm.weight", (dim,), "n"),
    ]
    return out


def weight_manifest(cfg: DualARConfig):
    """Every persistent tensor of the reference model: list of (key, shape, kind)."""
    m = [("embeddings.weight", (cfg.vocab_size, cfg.dim), "w"),
         ("codebook_embeddings.weight", (cfg.codebook_size * cfg.num_codebooks, cfg.dim), "w")]
    for i in range(cfg.n_layer):
        m += _layer_shapes(f"layers.{i}", cfg.dim, cfg.n_head, cfg.n_local_heads, cfg.head_dim,
                           cfg.intermediate_size, cfg.attention_qkv_bias, cfg.attention_o_bias,
                           cfg.attention_qk_norm)
    m.append(("norm.weight", (cfg.dim,), "n"))
    if not cfg.tie_word_embeddings:
        m.append(("output.weight", (cfg.vocab_size, cfg.dim), "w"))
    if cfg.fast_dim != cfg.dim:      # fast_project_in = nn.Linear(dim, fast_dim), registered before fast_embeddings (llama.py:510-516)
        m.append(("fast_project_in.weight", (cfg.fast_dim, cfg.dim), "w"))
        m.append(("fast_project_in.bias", (cfg.fast_dim,), "b"))
    m.append(("fast_embeddings.weight", (cfg.codebook_size, cfg.fast_dim), "w"))
    for i in range(cfg.n_fast_layer):
        m += _layer_shapes(f"fast_layers.{i}", cfg.fast_dim, cfg.fast_n_head, cfg.fast_n_local_heads,
                           cfg.fast_head_dim, cfg.fast_intermediate_size, cfg.fast_attention_qkv_bias,
                           cfg.fast_attention_o_bias, cfg.fast_attention_qk_norm)
    m.append(("fast_norm.weight", (cfg.fast_dim,), "n"))
    m.append(("fast_output.weight", (cfg.codebook_size, cfg.fast_dim), "w"))
    return m


def make_state_dict(cfg: DualARConfig, seed: int = 0, dtype=torch.bfloat16,
                    conditioned: bool = True, eos_reachable: bool = False) -> dict:
    """Seeded random-init weights on the CPU, keyed like the reference checkpoint.

    Linear / embedding weights ~ N(0, initializer_range) (llama.py:455-464); norm weights
    ~ 1 + 0.1 N(0,1) so a wrong norm-weight index cannot hide behind all-ones; biases small.
    ``eos_reachable`` flips the sign of the ``<|im_end|>`` head entry so EOS wins quickly
    (used by the early-stop tests).
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    std = cfg.initializer_range
    sd = {}
    for name, shape, kind in weight_manifest(cfg):
        t = torch.empty(shape, dtype=torch.float32)
        if kind == "w":
            t.normal_(0.0, std, generator=g)
        elif kind == "n":
            t.normal_(0.0, 0.1, generator=g).add_(1.0)
        else:
            t.normal_(0.0, 0.05, generator=g)
        sd[name] = t
    if conditioned:
        inter = cfg.intermediate_size
        units = min(CARRIER_UNITS, inter // 2)
        gain = CARRIER_GAIN * CARRIER_UNITS / units
        sd["layers.0.feed_forward.w3.weight"][:units] = sd["layers.0.feed_forward.w1.weight"][:units]
        sd["layers.0.feed_forward.w2.weight"][0, :units] = gain
        head = sd["embeddings.weight"] if cfg.tie_word_embeddings else sd["output.weight"]
        lo, hi = cfg.semantic_begin_id, cfg.semantic_end_id
        head[:lo] *= NONSEM_SCALE
        head[hi + 1:] *= NONSEM_SCALE
        head[:lo, 0] = NONSEM_BIAS
        head[hi + 1:, 0] = NONSEM_BIAS
        head[lo:hi + 1, 0] = 0.0
        if eos_reachable:
            head[cfg.im_end_id, 0] = -0.5 * NONSEM_BIAS
    return {k: v.to(dtype) for k, v in sd.items()}


def random_voice_codes(cfg: DualARConfig, n_frames: int, seed: int = 1) -> torch.Tensor:
    """A VoiceProfile-shaped code matrix (num_codebooks, n_frames) int64 (synthesizer.py:47-51):
    row 0 in [0, n_semantic), other rows in [0, min(1024, codebook_size))."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_sem = cfg.semantic_end_id - cfg.semantic_begin_id + 1
    fast_vocab = min(1024, cfg.codebook_size)
    top = torch.randint(0, n_sem, (1, n_frames), generator=g)
    rest = torch.randint(0, fast_vocab, (cfg.num_codebooks - 1, n_frames), generator=g)
    return torch.cat([top, rest], 0).to(torch.int64)


def synthetic_prompt(cfg: DualARConfig, n_text: int, n_frames: int, n_tail: int = 6,
                     seed: int = 1) -> torch.Tensor:
    """An encoded prompt ``(num_codebooks+1, T)`` int32 of the shape
    ``ContentSequence.encode_for_inference`` builds (inference.py:611-640): text ids with zero
    codebook rows, then a VQ span whose row 0 is ``semantic_begin_id + code0``, then a short
    text tail.  Text ids are drawn below the special-token range."""
    g = torch.Generator(device="cpu").manual_seed(seed + 1000)
    n_plain = min(cfg.semantic_begin_id, cfg.im_end_id) - 4
    T = n_text + n_frames + n_tail
    out = torch.zeros((cfg.codebook_dim, T), dtype=torch.int32)
    out[0, :n_text] = torch.randint(0, n_plain, (n_text,), generator=g).to(torch.int32)
    if n_frames:
        codes = random_voice_codes(cfg, n_frames, seed)
        out[0, n_text:n_text + n_frames] = (codes[0] + cfg.semantic_begin_id).to(torch.int32)
        out[1:, n_text:n_text + n_frames] = codes.to(torch.int32)
    out[0, n_text + n_frames:] = torch.randint(0, n_plain, (n_tail,), generator=g).to(torch.int32)
    return out
