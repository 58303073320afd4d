"""ORACLE SUPPORT (test infrastructure) -- run the UNMODIFIED reference from /root/reference.

Used to
  * validate ``oracle/dualar_oracle.py`` bit-for-bit against the real thing (build container),
  * generate the golden fixtures under ``tests/golden/`` (``tests/golden/make_golden.py``, build container), and
  * time the reference's own CPU and ``torch.compile`` paths beside ours in ``bench.py`` (GPU box: from the copy
    ``oracle/make_ref.py`` places under ``oracle/_ref/``, git-ignored, shipped with the snapshot).

What has to be faked so the reference imports and loads without network or vocoder deps
(SURVEY.md section 8c):
  * ``audiotools`` / ``dac`` are not installed and ``fish_tts/models/__init__.py:9`` imports the
    vocoder eagerly -> six symbols are stubbed in ``sys.modules`` before the import;
  * a model directory is fabricated: ``config.json``, a synthetic byte-level
    ``tokenizer.tiktoken`` with the right number of ranks so the special-token ids land where
    the real checkpoint puts them (tokenizer.py:84-101), optional ``special_tokens.json``,
    and ``model.pth`` = the seeded state dict of ``fish_tts_b200.synthetic``.
No reference source is copied; the package is imported from where it lies.
"""

from __future__ import annotations

import base64
import json
import os
import sys
import types
from pathlib import Path

import torch

def _find_root() -> Path:
    """/root/reference in the build container; oracle/_ref/ (the copy oracle/make_ref.py ships) on the GPU box"""
    cands = [Path(os.environ["FISH_TTS_REFERENCE"])] if os.environ.get("FISH_TTS_REFERENCE") else []
    cands += [Path("/root/reference"), Path(__file__).resolve().parent / "_ref"]
    for c in cands:
        if (c / "fish_tts" / "models" / "inference.py").exists():
            return c
    return cands[-1]


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return (REFERENCE_ROOT / "fish_tts" / "models" / "inference.py").exists()


def _install_stubs():
    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    import torch.nn as nn
    at, atml = mod("audiotools"), mod("audiotools.ml")
    at.ml = atml
    atml.BaseModel = type("BaseModel", (nn.Module,), {})
    dac, dm, dmb = mod("dac"), mod("dac.model"), mod("dac.model.base")
    dac.model, dm.base = dm, dmb
    dmb.CodecMixin = type("CodecMixin", (), {})
    dn, dnl, dnq = mod("dac.nn"), mod("dac.nn.layers"), mod("dac.nn.quantize")
    dac.nn, dn.layers, dn.quantize = dn, dnl, dnq
    dnl.Snake1d = type("Snake1d", (nn.Module,), {})
    dnl.WNConv1d = lambda *a, **k: nn.Conv1d(*a, **k)
    dnl.WNConvTranspose1d = lambda *a, **k: nn.ConvTranspose1d(*a, **k)
    dnq.ResidualVectorQuantize = type("ResidualVectorQuantize", (nn.Module,), {})


def import_reference():
    """Returns (fish_tts.models.llama, fish_tts.models.inference) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    import fish_tts.models.llama as llama
    import fish_tts.models.inference as inference
    return llama, inference


def n_ranks_for(cfg) -> int:
    """BPE rank count that puts <|im_end|> (special #4) at cfg.im_end_id."""
    return cfg.im_end_id - 4


def fabricate_model_dir(cfg, state_dict: dict, path: str | Path) -> Path:
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    with open(path / "config.json", "w") as f:
        json.dump(cfg.reference_json(), f)
    # synthetic byte-level vocabulary: 256 single bytes, then distinct 3-byte strings
    n = n_ranks_for(cfg)
    lines = []
    for r in range(n):
        tok = bytes([r]) if r < 256 else bytes([0xF0 | ((r >> 16) & 0x0F), (r >> 8) & 0xFF, r & 0xFF])
        lines.append(f"{base64.b64encode(tok).decode()} {r}")
    (path / "tokenizer.tiktoken").write_text("\n".join(lines) + "\n")
    n_sem = cfg.semantic_end_id - cfg.semantic_begin_id + 1
    if n_sem != 4096:
        # the 15 control tokens of tokenizer.py:52-66, then the semantic tokens
        specials = ["<|begin_of_text|>", "<|end_of_text|>", "<|pad|>", "<|im_start|>", "<|im_end|>", "<|phoneme_start|>",
                    "<|phoneme_end|>", "<|tool_call_start|>", "<|tool_call_end|>", "<|text|>", "<|voice|>", "<|interleave|>",
                    "<|audio_start|>", "<|audio_end|>", "<|audio|>"]
        specials += [f"<|semantic:{i}|>" for i in range(n_sem)]
        with open(path / "special_tokens.json", "w") as f:
            json.dump(specials, f)
    torch.save(state_dict, path / "model.pth")
    return path


def load_reference_model(cfg, state_dict: dict, workdir: str | Path, device="cpu", dtype=torch.bfloat16):
    """``init_model`` of the reference on a fabricated directory (inference.py:387-414), compile off."""
    _, inference = import_reference()
    d = fabricate_model_dir(cfg, state_dict, workdir)
    model, decode_one_token = inference.init_model(str(d), device=device, precision=dtype, compile=False)
    tok = model.tokenizer
    assert tok.semantic_begin_id == cfg.semantic_begin_id, (tok.semantic_begin_id, cfg.semantic_begin_id)
    assert tok.semantic_end_id == cfg.semantic_end_id
    assert tok.get_token_id("<|im_end|>") == cfg.im_end_id
    return model, decode_one_token


class RecordingStep:
    """Wraps the reference's ``decode_one_token_ar`` to capture per-step logits / tokens and to feed
    explicit noise, by patching the module-level hooks the reference already goes through
    (``inference.multinomial_sample_one_no_sync`` :24-27, ``model.forward_generate`` /
    ``forward_generate_fast``)."""

    def __init__(self, model, inference, noise_fn=None):
        self.model, self.inf, self.noise_fn = model, inference, noise_fn
        self.steps = []           # list of dict(slow_logits, hidden, fast_logits[], tokens)
        self._calls = 0

    def __enter__(self):
        inf, model = self.inf, self.model
        self._orig_mn = inf.multinomial_sample_one_no_sync
        self._orig_fg = model.forward_generate
        self._orig_ff = model.forward_generate_fast
        self._orig_step = inf.decode_one_token_ar
        rec = self

        def mn(probs_sort):
            if rec.noise_fn is None:
                return rec._orig_mn(probs_sort)
            q = rec.noise_fn(rec._calls, probs_sort.numel()).to(device=probs_sort.device, dtype=probs_sort.dtype)
            rec._calls += 1
            return torch.argmax(probs_sort / q, dim=-1, keepdim=True).to(dtype=torch.int)

        def fg(*a, **k):
            r = rec._orig_fg(*a, **k)
            rec._cur = {"slow_logits": r.logits[0, -1].clone(), "hidden": r.hidden_states[0, -1].clone(),
                        "fast_logits": []}
            return r

        def ff(x, input_pos=None):
            lg = rec._orig_ff(x, input_pos)
            if int(input_pos[0]) > 0:
                rec._cur["fast_logits"].append(lg[0, -1, :1024].clone())
            return lg

        def step(*a, **k):
            out = rec._orig_step(*a, **k)
            rec._cur["tokens"] = out[:, 0].clone()
            rec.steps.append(rec._cur)
            return out

        inf.multinomial_sample_one_no_sync = mn
        model.forward_generate = fg
        model.forward_generate_fast = ff
        inf.decode_one_token_ar = step
        self.step = step
        return self

    def __exit__(self, *exc):
        self.inf.multinomial_sample_one_no_sync = self._orig_mn
        self.inf.decode_one_token_ar = self._orig_step
        del self.model.forward_generate
        del self.model.forward_generate_fast
        return False
