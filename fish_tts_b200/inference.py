"""Host-side mirror of the reference's decode interface -- same names, argument meaning and error behaviour as
``fish_tts/models/inference.py``, with the B200 engine underneath (through the C-ABI, never through torch ops).

    reference                                   here
    ------------------------------------------  ------------------------------------------------------------
    init_model(path, device, precision, compile) -> (model, decode_one_token)      same signature  (:387-414)
    decode_one_token_ar(model, x, input_pos, T, p, rp, audio_masks, audio_parts, previous_tokens)  (:83-155)
    decode_n_tokens / decode_n_tokens_streaming                                    (:158-276)
    generate(*, model, prompt, max_new_tokens, audio_masks, audio_parts, decode_one_token, **sampling)  (:279-384)
    generate_streaming(...)                                                        (:643-738)

Two ways to use it:
  * stand-alone: ``init_model`` on a model directory (config.json + model.pth [+ tokenizer.tiktoken]);
  * dropped in beneath the UNMODIFIED reference: ``install()`` swaps the module attributes of
    ``fish_tts.models.inference`` that ``fish_tts/synthesizer.py`` looks up at call time (:161, 301, 454, 504), so
    ``FishTTS`` / ``get_instance`` / ``set_references`` / ``synthesize_stream`` run unchanged on top of the engine.
"""

from __future__ import annotations

import json
from pathlib import Path
from typing import Callable, Iterator, Optional

import numpy as np
import torch

from .config import DualARConfig
from .engine import DualAREngine

IM_END_TOKEN = "<|im_end|>"


class TokenIds:
    """What the hot path needs from FishTokenizer (tokenizer.py:84-101): three integers."""

    def __init__(self, semantic_begin_id: int, semantic_end_id: int, im_end_id: int):
        self.semantic_begin_id, self.semantic_end_id, self.im_end_id = semantic_begin_id, semantic_end_id, im_end_id

    def get_token_id(self, token: str) -> int:
        if token != IM_END_TOKEN:
            raise KeyError(token)
        return self.im_end_id

    @staticmethod
    def from_model_dir(path: Path) -> "TokenIds":
        """Special tokens follow the BPE ranks in enumeration order (tokenizer.py:84-87)."""
        n_ranks = sum(1 for line in open(path / "tokenizer.tiktoken") if line.strip())
        sp = path / "special_tokens.json"
        if sp.exists():
            specials = json.load(open(sp))
        else:   # ALL_SPECIAL_TOKENS, tokenizer.py:52-69: 15 control tokens then <|semantic:0..4095|>
            specials = ["<|begin_of_text|>", "<|end_of_text|>", "<|pad|>", "<|im_start|>", IM_END_TOKEN, "<|phoneme_start|>",
                        "<|phoneme_end|>", "<|tool_call_start|>", "<|tool_call_end|>", "<|text|>", "<|voice|>", "<|interleave|>",
                        "<|audio_start|>", "<|audio_end|>", "<|audio|>"] + [f"<|semantic:{i}|>" for i in range(4096)]
        ids = {t: n_ranks + i for i, t in enumerate(specials)}
        sem = sorted((int(t[len("<|semantic:"):-2]), i) for t, i in ids.items() if t.startswith("<|semantic:"))
        return TokenIds(sem[0][1], sem[-1][1], ids[IM_END_TOKEN])


class DualARModel:
    """Stands where the reference's ``DualARTransformer`` instance stands in the call chain: carries ``config``,
    ``tokenizer`` and the fixed sampling tensors the callers touch (inference.py:330-351, 400-404), and owns the engine."""

    def __init__(self, cfg: DualARConfig, engine: DualAREngine, tokenizer=None):
        self.config, self.engine = cfg, engine
        self.tokenizer = tokenizer or TokenIds(cfg.semantic_begin_id, cfg.semantic_end_id, cfg.im_end_id)
        dev = engine.device
        self.fixed_temperature = torch.tensor(0.7, device=dev, dtype=torch.float)
        self.fixed_top_p = torch.tensor(0.7, device=dev, dtype=torch.float)
        self.fixed_repetition_penalty = torch.tensor(1.5, device=dev, dtype=torch.float)
        self._cache_setup_done = True     # the engine allocates its KV caches at build time (llama.py:378-398)
        self.device = dev

    def setup_caches(self, max_batch_size: int = 1, max_seq_len: Optional[int] = None, dtype=torch.bfloat16):
        if max_batch_size != 1:
            raise ValueError("the decode path is batch 1, like the reference's sampler (inference.py:73)")

    def eval(self):
        return self

    def parameters(self):
        yield torch.empty(0, dtype=torch.bfloat16, device=self.device)


def _engine_of(model) -> DualAREngine:
    eng = getattr(model, "engine", None) or getattr(model, "_dualar_engine", None)
    if eng is None:
        raise RuntimeError("this model has no B200 engine attached; build it with fish_tts_b200.inference.init_model or install()")
    return eng


def init_model(checkpoint_path: str, device: str, precision: torch.dtype, compile: bool = False):
    """inference.py:387-414.  ``compile`` is accepted and ignored: the step is a CUDA graph either way."""
    if precision is not torch.bfloat16:
        raise ValueError("the B200 engine computes in bf16 (the FishTTS default precision, synthesizer.py:93)")
    if not str(device).startswith("cuda"):
        raise RuntimeError("the B200 decode path has no CPU fallback; use device='cuda'")
    path = Path(checkpoint_path)
    ids = TokenIds.from_model_dir(path)
    cfg = DualARConfig.from_json(path, semantic_begin_id=ids.semantic_begin_id, semantic_end_id=ids.semantic_end_id,
                                 im_end_id=ids.im_end_id)
    weights = torch.load(path / "model.pth", map_location="cpu", mmap=True, weights_only=True)   # llama.py:476-498
    if "state_dict" in weights:
        weights = weights["state_dict"]
    if next(iter(weights.keys())).startswith("model."):
        weights = {k.replace("model.", ""): v for k, v in weights.items()}
    weights = {k: v for k, v in weights.items() if "audio_" not in k}
    for k in [k for k in weights if k.endswith("attention.wq.weight")]:                            # legacy split qkv, llama.py:220-227
        pre = k[: -len("wq.weight")]
        weights[pre + "wqkv.weight"] = torch.cat([weights.pop(pre + "wq.weight"), weights.pop(pre + "wk.weight"), weights.pop(pre + "wv.weight")])
    dev = torch.device(device if ":" in str(device) else "cuda:0")
    engine = DualAREngine(cfg, weights, device=dev)
    model = DualARModel(cfg, engine, ids)
    return model, decode_one_token_ar


def decode_one_token_ar(model, x: torch.Tensor, input_pos: torch.Tensor, temperature: torch.Tensor, top_p: torch.Tensor,
                        repetition_penalty: torch.Tensor, audio_masks=None, audio_parts=None,
                        previous_tokens: Optional[torch.Tensor] = None) -> torch.Tensor:
    """inference.py:83-155.  x (1, C+1, T) int; T == 1 is the decode step, T > 1 the prefill call (:353-362).
    Returns (C+1, 1) int32 living in an engine-owned buffer (callers clone, :204)."""
    eng = _engine_of(model)
    if x.size(-1) == 1:
        return eng.step(x, input_pos, previous_tokens, temperature, top_p, repetition_penalty)
    if previous_tokens is not None:
        raise ValueError("a multi-position call is the prefill: the reference passes no previous_tokens there")
    eng.prefill(x.view(x.size(1), -1), 1, float(temperature), float(top_p), float(repetition_penalty))
    cols, _ = eng.collect()
    return torch.from_numpy(cols[:, :1].copy()).to(eng.device)


def decode_n_tokens(model, cur_token, input_pos, num_new_tokens, temperature, top_p, repetition_penalty,
                    audio_masks=None, audio_parts=None, decode_one_token: Callable = decode_one_token_ar, show_progress: bool = False):
    """inference.py:158-215, unchanged in structure (16-wide window, EOS break after recording the column)."""
    cfg = model.config
    codebook_dim = cfg.num_codebooks + 1
    previous_tokens = torch.zeros((codebook_dim, cfg.max_seq_len), dtype=torch.int, device=cur_token.device)
    end_token_id = model.tokenizer.get_token_id(IM_END_TOKEN)
    i = -1
    for i in range(num_new_tokens):
        window = previous_tokens[:, :16] if i < 16 else previous_tokens[:, i - 16: i]
        next_token = decode_one_token(model=model, x=cur_token, input_pos=input_pos, previous_tokens=window, temperature=temperature,
                                      top_p=top_p, repetition_penalty=repetition_penalty, audio_masks=audio_masks,
                                      audio_parts=audio_parts).clone()
        input_pos += 1
        cur_token = next_token.view(1, codebook_dim, -1)
        previous_tokens[:, i: i + 1] = next_token.view(codebook_dim, -1)
        if cur_token[0, 0, -1] == end_token_id:
            break
    return previous_tokens[:, : i + 1]


def _sampling(sampling_kwargs):
    return (sampling_kwargs.get("temperature", 0.7), sampling_kwargs.get("top_p", 0.7), sampling_kwargs.get("repetition_penalty", 1.5))


@torch.inference_mode()
def generate(*, model, prompt: torch.Tensor, max_new_tokens: int, audio_masks=None, audio_parts=None,
             decode_one_token: Callable = decode_one_token_ar, num_samples: int = 1, **sampling_kwargs) -> torch.Tensor:
    """inference.py:279-384 with the token loop on the device: one host round trip per 64 tokens instead of one per
    token, own prefill.  Returns (C+1, T + n) like the reference (the caller drops the last column, :839)."""
    eng = _engine_of(model)
    cfg = model.config
    T = prompt.size(1)
    if T >= cfg.max_seq_len:
        raise ValueError(f"Input sequence length {T} exceeds max_seq_len {cfg.max_seq_len}")
    t, p, rp = _sampling(sampling_kwargs)
    new = eng.generate(prompt.to(torch.int32).cpu().numpy(), max_new_tokens or 0, t, p, rp)
    seq = torch.empty((cfg.num_codebooks + 1, T + new.shape[1]), dtype=prompt.dtype, device=prompt.device)
    seq[:, :T] = prompt
    seq[:, T:] = torch.from_numpy(new).to(prompt.device)
    return seq


@torch.inference_mode()
def generate_streaming(*, model, prompt: torch.Tensor, max_new_tokens: int, audio_masks=None, audio_parts=None,
                       decode_one_token: Callable = decode_one_token_ar, first_chunk: int = 10, chunk: int = 20,
                       **sampling_kwargs) -> Iterator[torch.Tensor]:
    """inference.py:643-738: yields (C, 1) code columns as they are produced (the semantic row is dropped, :721, 271).
    Columns leave the device in chunks (``first_chunk`` then ``chunk`` columns, the sizes ``synthesize_stream`` batches for the
    vocoder, synthesizer.py:548-559) through pinned host buffers with an event per chunk; the next chunk's decode steps are
    already enqueued while this one is consumed (``DualAREngine.stream``)."""
    eng = _engine_of(model)
    cfg = model.config
    T = prompt.size(1)
    if T >= cfg.max_seq_len:
        raise ValueError(f"Input sequence length {T} exceeds max_seq_len {cfg.max_seq_len}")
    if not max_new_tokens or T + max_new_tokens > cfg.max_seq_len:
        max_new_tokens = cfg.max_seq_len - T
    t, p, rp = _sampling(sampling_kwargs)
    for cols in eng.stream(prompt.to(torch.int32).cpu().numpy(), max_new_tokens, t, p, rp, first_chunk=first_chunk, chunk=chunk):
        block = torch.from_numpy(cols[1:]).to(prompt.device)
        for j in range(block.size(1)):
            yield block[:, j: j + 1]


# ---- drop-in beneath the unmodified reference -------------------------------------------------------------------------
def install(inference_module=None):
    """Swap the reference's decode path for the engine without editing any reference file.

    ``fish_tts/synthesizer.py`` imports ``init_model`` / ``generate_long`` from ``fish_tts.models.inference`` at CALL time
    (:161, 301, 454, 504) and ``generate_long`` looks up ``generate`` / ``generate_streaming`` as module globals (:805,
    824), so replacing these module attributes is enough; the reference's own model object (weights, tokenizer, prompt
    builder) stays in charge of everything outside the hot path.  Returns the module that was patched."""
    if inference_module is None:
        import fish_tts.models.inference as inference_module
    ref = inference_module
    if getattr(ref, "_dualar_installed", False):
        return ref
    orig_init = ref.init_model

    def init_model_b200(checkpoint_path: str, device: str, precision: torch.dtype, compile: bool = False):
        model, _ = orig_init(checkpoint_path, device, precision, compile=False)      # reference loads weights + tokenizer
        tok = model.tokenizer
        cfg_fields = {f: getattr(model.config, f) for f in DualARConfig.__dataclass_fields__ if hasattr(model.config, f)}
        cfg = DualARConfig(**cfg_fields)
        cfg.semantic_begin_id, cfg.semantic_end_id = tok.semantic_begin_id, tok.semantic_end_id
        cfg.im_end_id = tok.get_token_id(IM_END_TOKEN)
        sd = {k: v for k, v in model.state_dict().items() if not k.endswith("_cache")}
        dev = next(model.parameters()).device
        model._dualar_engine = DualAREngine(cfg, sd, device=dev, freqs_cis=model.freqs_cis, fast_freqs_cis=model.fast_freqs_cis)
        model._dualar_config = cfg
        # the engine holds its own repacked copy of every weight: release the reference module's (1.4 GB for s1-mini, plus the
        # max_seq_len^2 causal mask) instead of keeping both on the GPU.  The module stays in charge of config / tokenizer / prompt
        # building only; its forward is never called again (generate / generate_streaming / decode_one_token_ar are replaced).
        del sd
        for prm in model.parameters():
            prm.data = torch.empty(0, device=dev, dtype=prm.dtype)
        for name, buf in list(model.named_buffers()):
            if buf.numel() > 1024:
                mod, _, leaf = name.rpartition(".")
                setattr(model.get_submodule(mod) if mod else model, leaf, torch.empty(0, device=dev, dtype=buf.dtype))
        if dev.type == "cuda":
            torch.cuda.empty_cache()
        return model, decode_one_token_ar

    def _cfg_model(model):
        class _View:      # the reference model with the engine's config view
            config, tokenizer, engine = model._dualar_config, model.tokenizer, model._dualar_engine
        return _View

    def generate_b200(*, model, **kw):
        return generate(model=_cfg_model(model), **kw)

    def generate_streaming_b200(*, model, **kw):
        return generate_streaming(model=_cfg_model(model), **kw)

    ref._dualar_original = {n: getattr(ref, n) for n in ("init_model", "generate", "generate_streaming", "decode_one_token_ar")}
    ref.init_model = init_model_b200
    ref.generate = torch.inference_mode()(generate_b200)
    ref.generate_streaming = generate_streaming_b200
    ref.decode_one_token_ar = decode_one_token_ar
    ref._dualar_installed = True
    return ref


def uninstall(inference_module=None):
    if inference_module is None:
        import fish_tts.models.inference as inference_module
    for n, f in getattr(inference_module, "_dualar_original", {}).items():
        setattr(inference_module, n, f)
    inference_module._dualar_installed = False
