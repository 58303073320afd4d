"""ORACLE (test infrastructure, not product code) -- a functional restatement of the
reference's dual-AR decode path in plain torch ops.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
legs may import this module.  The product path (``fish_tts_b200``) never does: it calls
the CUDA extension through the C-ABI and fails loudly when that is missing.

What it restates (all citations are ``/root/reference/`` file:line):
  * ``DualARTransformer`` inference forward  -- fish_tts/models/llama.py:400-453, 561-591
  * ``Attention`` / ``KVCache`` / ``RMSNorm`` / ``FeedForward`` / RoPE -- llama.py:126-331, 594-618
  * ``sample`` / ``logits_to_probs`` / ``multinomial_sample_one_no_sync`` -- inference.py:24-80
  * ``decode_one_token_ar`` -- inference.py:83-155
  * ``decode_n_tokens`` / ``generate`` (prefill + loop + EOS) -- inference.py:158-215, 279-384

It is written against a flat ``{state_dict key: tensor}`` dict instead of ``nn.Module``s, and
issues the *same* ATen ops in the same order as the reference, so on one device and one
torch build its outputs are bit-identical to the reference's eager path.

PINNING: the reference ships no golden vectors for this path.  The oracle
is pinned instead against the reference itself: ``tests/golden/make_golden.py`` imports the
unmodified reference from ``/root/reference`` (CPU), runs it on the seeded checkpoints of
``fish_tts_b200.synthetic`` and commits its per-step logits / tokens under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those, and
``tests/test_oracle_vs_reference.py`` re-runs the comparison live whenever ``/root/reference``
is present.

Two knobs exist only because the reference's behaviour is implementation-defined there, and
both default to the reference's own calls:
  * ``noise`` -- the Exp(1) draws of ``multinomial_sample_one_no_sync`` can be supplied
    explicitly (shared-RNG parity); default draws them with ``exponential_`` like the reference.
  * ``stable_ties`` -- ``torch.sort`` leaves the order of equal logits unspecified; ``True``
    asks for index order (what the CUDA sampler implements) and computes the nucleus cumsum as
    "fp32 running sum, rounded to bf16 per element" (the reference's CPU semantics) on every device.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional

import torch
import torch.nn.functional as F
from torch import Tensor
from torch.nn.attention import SDPBackend, sdpa_kernel


# --------------------------------------------------------------------------------------
# model container
# --------------------------------------------------------------------------------------

def precompute_freqs_cis(seq_len: int, n_elem: int, base: float = 10000) -> Tensor:
    """llama.py:594-603 -- (seq_len, n_elem/2, 2) cos/sin table, stored in bf16."""
    freqs = 1.0 / (base ** (torch.arange(0, n_elem, 2)[: (n_elem // 2)].float() / n_elem))
    t = torch.arange(seq_len, device=freqs.device)
    freqs = torch.outer(t, freqs)
    freqs_cis = torch.polar(torch.ones_like(freqs), freqs)
    cache = torch.stack([freqs_cis.real, freqs_cis.imag], dim=-1)
    return cache.to(dtype=torch.bfloat16)


@dataclass
class OracleModel:
    """Weights + KV caches of one DualARTransformer instance (llama.py:334-376, 503-559)."""
    cfg: object                      # fish_tts_b200.config.DualARConfig
    w: dict                          # state_dict-keyed tensors on `device`
    device: torch.device
    dtype: torch.dtype
    freqs_cis: Tensor = None
    fast_freqs_cis: Tensor = None
    accum: Optional[torch.dtype] = None           # None: the reference's own F.linear (cuBLAS / MKL accumulation order).  torch.float64:
                                                  # "ground truth" -- every linear accumulated in fp64 and rounded to the model dtype at the
                                                  # same point, i.e. the reference's rounding points WITHOUT any fp32 summation-order noise
    kv: list = field(default_factory=list)        # per slow layer [k_cache, v_cache]
    fast_kv: list = field(default_factory=list)   # per fast layer [k_cache, v_cache]
    max_seq_len: int = -1

    @staticmethod
    def build(cfg, state_dict: dict, device="cpu", dtype=torch.bfloat16) -> "OracleModel":
        device = torch.device(device)
        w = {k: v.to(device=device, dtype=dtype) for k, v in state_dict.items()}
        m = OracleModel(cfg=cfg, w=w, device=device, dtype=dtype)
        # non-persistent buffers (llama.py:361-370, 537-541); `.to(dtype)` of the module casts them
        m.freqs_cis = precompute_freqs_cis(cfg.max_seq_len, cfg.head_dim, cfg.rope_base).to(device=device, dtype=dtype)
        m.fast_freqs_cis = precompute_freqs_cis(cfg.num_codebooks, cfg.fast_head_dim, cfg.rope_base).to(device=device, dtype=dtype)
        return m

    def setup_caches(self, max_seq_len: Optional[int] = None):
        """llama.py:378-398, 544-559 (batch 1)."""
        cfg = self.cfg
        max_seq_len = max_seq_len or cfg.max_seq_len
        if self.max_seq_len >= max_seq_len:
            return
        max_seq_len = max_seq_len if max_seq_len % 8 == 0 else max_seq_len + 8 - max_seq_len % 8
        self.max_seq_len = max_seq_len
        z = lambda h, s, d: torch.zeros((1, h, s, d), dtype=self.dtype, device=self.device)
        self.kv = [[z(cfg.n_local_heads, max_seq_len, cfg.head_dim) for _ in range(2)]
                   for _ in range(cfg.n_layer)]
        self.fast_kv = [[z(cfg.fast_n_local_heads, cfg.num_codebooks, cfg.fast_head_dim) for _ in range(2)]
                        for _ in range(cfg.n_fast_layer)]

    def causal_rows(self, input_pos: Tensor, width: int) -> Tensor:
        """``causal_mask[None, None, input_pos, :width]`` (llama.py:366-370, 437, 569-571) without
        materialising the (S, S) table: row t is True for columns <= input_pos[t]."""
        cols = torch.arange(width, device=self.device)
        return (cols[None, :] <= input_pos.to(torch.long)[:, None])[None, None]


# --------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------

def linear(m: "OracleModel", x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    """nn.Linear as the reference calls it (m.accum is None), or the same contraction accumulated in m.accum (tests only: the
    yardstick against which both our kernels and torch's own bf16 GEMMs are measured, tests/test_gpu_parity.py)."""
    if m.accum is None:
        return F.linear(x, w, b)
    y = F.linear(x.to(m.accum), w.to(m.accum), None if b is None else b.to(m.accum))
    return y.to(x.dtype)


def rms_norm(x: Tensor, weight: Tensor, eps: float) -> Tensor:
    """llama.py:172-177 -- fp32 normalise, round to x.dtype, THEN multiply by weight."""
    xf = x.float()
    out = (xf * torch.rsqrt(torch.mean(xf * xf, dim=-1, keepdim=True) + eps)).type_as(x)
    return out * weight


def apply_rotary_emb(x: Tensor, freqs_cis: Tensor) -> Tensor:
    """llama.py:606-618 -- interleaved-pair rotation, fp32 math, table in model dtype."""
    xshaped = x.float().reshape(*x.shape[:-1], -1, 2)
    freqs_cis = freqs_cis.view(1, xshaped.size(1), 1, xshaped.size(3), 2)
    x_out2 = torch.stack(
        [xshaped[..., 0] * freqs_cis[..., 0] - xshaped[..., 1] * freqs_cis[..., 1],
         xshaped[..., 1] * freqs_cis[..., 0] + xshaped[..., 0] * freqs_cis[..., 1]], -1)
    return x_out2.flatten(3).type_as(x)


def eq_scaled_dot_product_attention(query, key, value, attn_mask) -> Tensor:
    """llama.py:285-309 -- the fast layers' hand-written attention (model dtype throughout)."""
    L, S = query.size(-2), key.size(-2)
    scale_factor = 1 / math.sqrt(query.size(-1))
    attn_bias = torch.zeros(1, 1, L, S, dtype=query.dtype, device=query.device)
    attn_bias.masked_fill_(attn_mask.logical_not(), float("-inf"))
    attn_weight = query @ key.transpose(-2, -1) * scale_factor
    attn_weight += attn_bias
    attn_weight = torch.softmax(attn_weight, dim=-1)
    attn_weight = torch.dropout(attn_weight, 0.0, train=True)
    return attn_weight @ value


def attention(m: OracleModel, prefix: str, kv, x, freqs_cis, mask, input_pos,
              n_head, n_local_heads, head_dim, qk_norm, use_sdpa) -> Tensor:
    """llama.py:229-283."""
    w = m.w
    bsz, seqlen, _ = x.shape
    q_size, kv_size = n_head * head_dim, n_local_heads * head_dim
    qkv = linear(m, x, w[f"{prefix}.wqkv.weight"], w.get(f"{prefix}.wqkv.bias"))
    q, k, v = qkv.split([q_size, kv_size, kv_size], dim=-1)
    q = q.view(bsz, seqlen, n_head, head_dim)
    k = k.view(bsz, seqlen, n_local_heads, head_dim)
    v = v.view(bsz, seqlen, n_local_heads, head_dim)
    if qk_norm:
        q = F.rms_norm(q, (head_dim,), w[f"{prefix}.q_norm.weight"], m.cfg.norm_eps)
        k = F.rms_norm(k, (head_dim,), w[f"{prefix}.k_norm.weight"], m.cfg.norm_eps)
    q = apply_rotary_emb(q, freqs_cis)
    k = apply_rotary_emb(k, freqs_cis)
    q, k, v = map(lambda t: t.transpose(1, 2), (q, k, v))
    # KVCache.update, llama.py:142-149
    kv[0][:, :, input_pos] = k
    kv[1][:, :, input_pos] = v
    k, v = kv[0], kv[1]
    k = k.repeat_interleave(n_head // n_local_heads, dim=1)
    v = v.repeat_interleave(n_head // n_local_heads, dim=1)
    if use_sdpa:
        y = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0)
    else:
        y = eq_scaled_dot_product_attention(q, k, v, mask)
    y = y.transpose(1, 2).contiguous().view(bsz, seqlen, q_size)
    return linear(m, y, w[f"{prefix}.wo.weight"], w.get(f"{prefix}.wo.bias"))


def block(m: OracleModel, prefix: str, kv, x, freqs_cis, mask, input_pos, fast: bool) -> Tensor:
    """llama.py:322-331 + 189-190."""
    cfg, w = m.cfg, m.w
    if fast:
        dims = (cfg.fast_n_head, cfg.fast_n_local_heads, cfg.fast_head_dim, cfg.fast_attention_qk_norm, False)
    else:
        dims = (cfg.n_head, cfg.n_local_heads, cfg.head_dim, cfg.attention_qk_norm, True)
    h = x + attention(m, f"{prefix}.attention", kv, rms_norm(x, w[f"{prefix}.attention_norm.weight"], cfg.norm_eps),
                      freqs_cis, mask, input_pos, *dims)
    hn = rms_norm(h, w[f"{prefix}.ffn_norm.weight"], cfg.norm_eps)
    ff = linear(m, F.silu(linear(m, hn, w[f"{prefix}.feed_forward.w1.weight"])) *
                linear(m, hn, w[f"{prefix}.feed_forward.w3.weight"]),
                w[f"{prefix}.feed_forward.w2.weight"])
    return h + ff


def forward_generate(m: OracleModel, inp: Tensor, input_pos: Tensor):
    """llama.py:400-453 (+ override 582-591: hidden_states = fast_project_in(x), an nn.Linear(dim, fast_dim) when fast_dim != dim,
    llama.py:510-513, Identity otherwise).
    inp: (1, C+1, T) int.  Returns (logits (1,1,V), hidden_states (1,1,dim) -- the UN-normalised x)."""
    cfg, w = m.cfg, m.w
    embeds = []
    for i in range(cfg.num_codebooks):
        embeds.append(F.embedding(inp[:, i + 1] + i * cfg.codebook_size, w["codebook_embeddings.weight"]))
    vq_embeds_sum = torch.stack(embeds, dim=1).sum(dim=1)
    vq_masks = (inp[:, 0] >= cfg.semantic_begin_id) & (inp[:, 0] <= cfg.semantic_end_id)
    vq_embeds_sum[~vq_masks] = 0
    x = F.embedding(inp[:, 0], w["embeddings.weight"]) + vq_embeds_sum
    if cfg.scale_codebook_embeddings:
        vq_masks_expanded = vq_masks.unsqueeze(-1).expand_as(x)
        x = torch.where(vq_masks_expanded, x / math.sqrt(cfg.num_codebooks + 1), x)
    mask = m.causal_rows(input_pos, m.max_seq_len)
    freqs_cis = m.freqs_cis[input_pos]
    for i in range(cfg.n_layer):
        x = block(m, f"layers.{i}", m.kv[i], x, freqs_cis, mask, input_pos, fast=False)
    if x.size(1) > 1:
        x = x[:, -1:]
    slow_out = rms_norm(x, w["norm.weight"], cfg.norm_eps)
    head = w["embeddings.weight"] if cfg.tie_word_embeddings else w["output.weight"]
    logits = linear(m, slow_out, head)
    if "fast_project_in.weight" in w:      # DualARTransformer.forward_generate, llama.py:590
        x = linear(m, x, w["fast_project_in.weight"], w.get("fast_project_in.bias"))
    return logits, x


def forward_generate_fast(m: OracleModel, x: Tensor, input_pos: Tensor) -> Tensor:
    """llama.py:561-580."""
    cfg, w = m.cfg, m.w
    x = x.view(x.shape[0], 1, -1)
    fast_mask = m.causal_rows(input_pos, cfg.num_codebooks)
    fast_freqs_cis = m.fast_freqs_cis[input_pos]
    for i in range(cfg.n_fast_layer):
        x = block(m, f"fast_layers.{i}", m.fast_kv[i], x, fast_freqs_cis, fast_mask, input_pos, fast=True)
    fast_out = rms_norm(x, w["fast_norm.weight"], cfg.norm_eps)
    return linear(m, fast_out, w["fast_output.weight"])


# --------------------------------------------------------------------------------------
# sampling
# --------------------------------------------------------------------------------------

class NoiseSource:
    """Explicit Exp(1) draws for ``multinomial_sample_one_no_sync``.  ``next(n)`` returns the
    noise vector for the next head (slow head first, then fast heads 1..C-1, every step)."""

    def __init__(self, fn: Callable[[int, int], Tensor]):
        self.fn, self.calls = fn, 0

    def next(self, n: int) -> Tensor:
        q = self.fn(self.calls, n)
        self.calls += 1
        return q


def multinomial_sample_one_no_sync(probs_sort: Tensor, noise: Optional[NoiseSource]) -> Tensor:
    """inference.py:24-27."""
    if noise is None:
        q = torch.empty_like(probs_sort).exponential_(1)
    else:
        q = noise.next(probs_sort.numel()).to(device=probs_sort.device, dtype=probs_sort.dtype)
    return torch.argmax(probs_sort / q, dim=-1, keepdim=True).to(dtype=torch.int)


def logits_to_probs(logits, temperature, top_p, repetition_penalty, previous_tokens=None,
                    stable_ties: bool = False) -> Tensor:
    """inference.py:30-61 (logits is the 1-D view ``logits[0, -1]``, modified in place)."""
    if previous_tokens is not None:
        previous_tokens = previous_tokens.long()
        score = torch.gather(logits, dim=-1, index=previous_tokens)
        score = torch.where(score < 0, score * repetition_penalty, score / repetition_penalty)
        logits.scatter_(dim=-1, index=previous_tokens, src=score)
    sorted_logits, sorted_indices = torch.sort(logits, descending=True, stable=stable_ties)
    probs_sorted = torch.nn.functional.softmax(sorted_logits, dim=-1)
    if stable_ties:
        # the reference's CPU cumsum: fp32 running sum, each element rounded to logits.dtype
        cum_probs = torch.cumsum(probs_sorted.float(), dim=-1).to(probs_sorted.dtype)
    else:
        cum_probs = torch.cumsum(probs_sorted, dim=-1)
    sorted_indices_to_remove = cum_probs > top_p
    sorted_indices_to_remove[0] = False
    indices_to_remove = sorted_indices_to_remove.scatter(dim=-1, index=sorted_indices, src=sorted_indices_to_remove)
    logits = logits.masked_fill(indices_to_remove, -float("Inf"))
    logits = logits / torch.clip(temperature, min=1e-5)
    return torch.nn.functional.softmax(logits, dim=-1)


def sample(logits, temperature, top_p, repetition_penalty, previous_tokens=None,
           noise: Optional[NoiseSource] = None, stable_ties: bool = False):
    """inference.py:64-80."""
    probs = logits_to_probs(logits[0, -1], temperature, top_p, repetition_penalty, previous_tokens, stable_ties)
    return multinomial_sample_one_no_sync(probs, noise), probs


# --------------------------------------------------------------------------------------
# the decode step and the loops around it
# --------------------------------------------------------------------------------------

@dataclass
class StepTrace:
    """What a parity test wants to see from one step (filled only when a list is passed)."""
    slow_logits: Tensor = None      # (V,) BEFORE the in-place repetition penalty
    hidden: Tensor = None           # (dim,) un-normalised last-layer output
    fast_logits: list = field(default_factory=list)   # per codebook k>=1: (fast_vocab,) before penalty
    tokens: Tensor = None           # (C+1,) int32


def decode_one_token_ar(m: OracleModel, x: Tensor, input_pos: Tensor, temperature: Tensor, top_p: Tensor,
                        repetition_penalty: Tensor, previous_tokens: Optional[Tensor] = None,
                        noise: Optional[NoiseSource] = None, stable_ties: bool = False,
                        trace: Optional[list] = None) -> Tensor:
    """inference.py:83-155.  x: (1, C+1, T) int; returns (C+1, 1) int32."""
    cfg = m.cfg
    logits, hidden_states = forward_generate(m, x, input_pos)
    tr = StepTrace() if trace is not None else None
    if tr is not None:
        tr.slow_logits = logits[0, -1].clone()
        tr.hidden = hidden_states[0, -1].clone()
    codebooks = [sample(logits, temperature, top_p, repetition_penalty,
                        previous_tokens[:, 0] if previous_tokens is not None else None,
                        noise, stable_ties)[0]]
    for kv in m.fast_kv:           # inference.py:115-119
        kv[0].fill_(0)
        kv[1].fill_(0)
    input_pos = torch.tensor([0], device=hidden_states.device, dtype=torch.long)
    forward_generate_fast(m, hidden_states, input_pos)
    a = codebooks[0] - cfg.semantic_begin_id
    a[a < 0] = 0
    hidden_states = F.embedding(a, m.w["fast_embeddings.weight"])
    codebooks.append(a)
    for codebook_idx in range(1, cfg.num_codebooks):
        input_pos = torch.tensor([codebook_idx], device=hidden_states.device, dtype=torch.long)
        logits = forward_generate_fast(m, hidden_states, input_pos)
        short_logits = logits[:, :, :1024]
        if tr is not None:
            tr.fast_logits.append(short_logits[0, -1].clone())
        a = sample(short_logits, temperature, top_p, repetition_penalty,
                   previous_tokens[codebook_idx + 1] if previous_tokens is not None else None,
                   noise, stable_ties)[0]
        hidden_states = F.embedding(a, m.w["fast_embeddings.weight"])
        codebooks.append(a)
    out = torch.stack(codebooks, dim=1).T
    if tr is not None:
        tr.tokens = out[:, 0].clone()
        trace.append(tr)
    return out


def decode_n_tokens(m: OracleModel, cur_token, input_pos, num_new_tokens, temperature, top_p,
                    repetition_penalty, noise=None, stable_ties=False, trace=None,
                    decode_one_token: Optional[Callable] = None):
    """inference.py:158-215 (the 16-wide window, the MATH sdpa context, the EOS break)."""
    cfg = m.cfg
    codebook_dim = cfg.num_codebooks + 1
    previous_tokens = torch.zeros((codebook_dim, cfg.max_seq_len), dtype=torch.int, device=cur_token.device)
    i = -1
    for i in range(num_new_tokens):
        win_size = 16
        window = previous_tokens[:, :win_size] if i < win_size else previous_tokens[:, i - win_size: i]
        with sdpa_kernel(SDPBackend.MATH):
            if decode_one_token is None:
                next_token = decode_one_token_ar(m, cur_token, input_pos, temperature, top_p, repetition_penalty,
                                                 window, noise, stable_ties, trace).clone()
            else:
                next_token = decode_one_token(x=cur_token, input_pos=input_pos, previous_tokens=window,
                                              temperature=temperature, top_p=top_p,
                                              repetition_penalty=repetition_penalty).clone()
        input_pos += 1
        cur_token = next_token.view(1, codebook_dim, -1)
        previous_tokens[:, i: i + 1] = next_token.view(codebook_dim, -1)
        if cur_token[0, 0, -1] == cfg.im_end_id:
            break
    return previous_tokens[:, : i + 1]


@torch.inference_mode()
def generate(m: OracleModel, prompt: Tensor, max_new_tokens: int, temperature: float = 0.7, top_p: float = 0.7,
             repetition_penalty: float = 1.5, noise=None, stable_ties=False, trace=None,
             math_prefill: bool = False, decode_one_token: Optional[Callable] = None) -> Tensor:
    """inference.py:279-384.  prompt (C+1, T) int32 on m.device -> (C+1, T + n) int32.

    ``math_prefill=True`` runs the prefill under the MATH sdpa backend too (the reference leaves the
    prefill to the default backend, inference.py:353); the CUDA engine's prefill uses the math semantics.
    """
    cfg = m.cfg
    T = prompt.size(1)
    if T >= cfg.max_seq_len:
        raise ValueError(f"Input sequence length {T} exceeds max_seq_len {cfg.max_seq_len}")
    if max_new_tokens:
        if T + max_new_tokens > cfg.max_seq_len:
            max_new_tokens = cfg.max_seq_len - T
    else:
        max_new_tokens = cfg.max_seq_len - T
    device = prompt.device
    m.setup_caches(cfg.max_seq_len)
    codebook_dim = 1 + cfg.num_codebooks
    input_pos = torch.arange(0, T, device=device, dtype=torch.long)
    seq = torch.empty((codebook_dim, cfg.max_seq_len), dtype=prompt.dtype, device=device)
    seq[:, :T] = prompt
    t_ = torch.tensor(temperature, device=device, dtype=torch.float)
    p_ = torch.tensor(top_p, device=device, dtype=torch.float)
    r_ = torch.tensor(repetition_penalty, device=device, dtype=torch.float)
    if math_prefill:
        with sdpa_kernel(SDPBackend.MATH):
            first_token = decode_one_token_ar(m, prompt.view(1, codebook_dim, -1), input_pos, t_, p_, r_,
                                              None, noise, stable_ties, trace)
    else:
        first_token = decode_one_token_ar(m, prompt.view(1, codebook_dim, -1), input_pos, t_, p_, r_,
                                          None, noise, stable_ties, trace)
    seq[:, T: T + 1] = first_token
    input_pos = torch.tensor([T], device=device, dtype=torch.int)
    x = decode_n_tokens(m, first_token.view(1, codebook_dim, -1), input_pos, max_new_tokens - 1, t_, p_, r_,
                        noise, stable_ties, trace, decode_one_token)
    seq = seq[:, : T + 1 + x.size(1)]
    seq[:, T + 1:] = x
    return seq
