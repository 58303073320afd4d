"""FishTTS / VoiceProfile / get_instance / reset_instance -- the reference's public API (fish_tts/synthesizer.py:47-719)
kept name for name, with the dual-AR decode running on the B200 engine.

Scope: the decode path is ours; the codec / vocoder and the BPE tokenizer are NOT -- the reference keeps them
on its torch path and they are un-vendored there (``dac``, ``audiotools``, ``tiktoken``).  So:
  * when the reference package is importable, the simplest drop-in is ``fish_tts_b200.inference.install()`` followed by the
    reference's own ``get_instance()`` -- its ``FishTTS`` then runs unchanged on the engine;
  * this stand-alone class needs the two out-of-scope pieces handed in: ``prompt_encoder(texts, codes, text) -> (C+1, T) int32``
    (what ``ContentSequence.encode_for_inference`` builds, inference.py:611-640) and ``vocoder(codes (C, n) int) -> float audio``.
    Without a vocoder ``synthesize`` raises ``RuntimeError("Vocoder not loaded")`` exactly like the reference (:599-600).
"""

from __future__ import annotations

import io
import logging
import queue
import threading
import wave
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, Iterator, Literal, Optional

import numpy as np
import torch

from . import inference

logger = logging.getLogger(__name__)

_instance: "FishTTS | None" = None
_instance_lock = threading.Lock()


@dataclass
class VoiceProfile:
    """Encoded reference audio: ``codes`` (num_codebooks, seq_len) int64 + transcript (synthesizer.py:47-65)."""
    codes: np.ndarray
    text: str = ""
    name: str = ""

    def save(self, path) -> None:
        np.save(path, self.codes)

    @classmethod
    def load(cls, path, text: str = "", name: str = "") -> "VoiceProfile":
        return cls(codes=np.load(path), text=text, name=name or Path(path).stem)


@dataclass
class _PrefillCache:
    prompt_text: list = field(default_factory=list)
    prompt_tokens: list = field(default_factory=list)
    profiles: list = field(default_factory=list)


class FishTTS:
    def __init__(self, model_dir, device: Literal["cpu", "cuda"] = "cuda", precision: Literal["bf16", "fp16", "fp32"] = "bf16",
                 warmup: bool = True, prompt_encoder: Optional[Callable] = None, vocoder: Optional[Callable] = None):
        if device != "cuda" or precision != "bf16":
            raise RuntimeError("the B200 decode path runs on cuda in bf16 only (no CPU fallback)")
        if model_dir is None:
            raise RuntimeError("no network here: pass model_dir (the reference downloads fishaudio/openaudio-s1-mini, synthesizer.py:146-157)")
        self.device, self._precision = device, precision
        self._prefill_cache, self._prefill_lock = _PrefillCache(), threading.Lock()
        self._model, self._decode_one_token = inference.init_model(str(model_dir), device, torch.bfloat16, compile=True)
        self._prompt_encoder, self._vocoder = prompt_encoder, vocoder
        self._is_warmed_up = False
        if warmup:   # the reference's warmup triggers Inductor (synthesizer.py:295-323); ours just touches the graphs
            cfg = self._model.config
            p = np.zeros((cfg.num_codebooks + 1, 4), dtype=np.int32)
            self._model.engine.generate(p, 4, 0.7, 0.8, 1.1)
            self._is_warmed_up = True

    # ---- reference management (synthesizer.py:363-429) ---------------------------------------------------------------------
    def set_references(self, profiles: list) -> None:
        with self._prefill_lock:
            self._prefill_cache = _PrefillCache([p.text for p in profiles], [torch.from_numpy(p.codes) for p in profiles], list(profiles))

    def add_reference(self, profile: VoiceProfile) -> None:
        with self._prefill_lock:
            self._prefill_cache.profiles.append(profile)
            self._prefill_cache.prompt_text.append(profile.text)
            self._prefill_cache.prompt_tokens.append(torch.from_numpy(profile.codes))

    def clear_references(self) -> None:
        with self._prefill_lock:
            self._prefill_cache = _PrefillCache()

    def get_references(self) -> list:
        with self._prefill_lock:
            return list(self._prefill_cache.profiles)

    @property
    def num_references(self) -> int:
        return len(self._prefill_cache.profiles)

    def _get_prompt_data(self, references):
        if references is not None:
            return [p.text for p in references], [torch.from_numpy(p.codes) for p in references]
        with self._prefill_lock:
            return list(self._prefill_cache.prompt_text), list(self._prefill_cache.prompt_tokens)

    def _encode_prompt(self, text, references) -> torch.Tensor:
        if self._prompt_encoder is None:
            raise RuntimeError("no prompt_encoder: text -> ids needs the reference's tokenizer / ContentSequence (out of scope here)")
        texts, codes = self._get_prompt_data(references)
        prompt = torch.as_tensor(self._prompt_encoder(texts, codes, text), dtype=torch.int32)
        cfg = self._model.config
        if prompt.size(1) > cfg.max_seq_len - 2048:
            raise ValueError(f"Prompt is too long: {prompt.size(1)} > {cfg.max_seq_len - 2048}")      # inference.py:794-795
        return prompt

    # ---- generation ---------------------------------------------------------------------------------------------------------
    def generate_codes(self, text: str, references=None, temperature=0.7, top_p=0.8, repetition_penalty=1.1, max_tokens=2048) -> torch.Tensor:
        """codes (C, n): what the reference hands to the vocoder (``y[1:, prompt_length:-1]``, inference.py:839)"""
        assert 0 < top_p <= 1 and 0 < repetition_penalty < 2 and 0 < temperature < 2                    # inference.py:763-765
        prompt = self._encode_prompt(text, references)
        y = inference.generate(model=self._model, prompt=prompt, max_new_tokens=max_tokens, temperature=temperature, top_p=top_p,
                               repetition_penalty=repetition_penalty)
        return y[1:, prompt.size(1):-1].clone()

    def synthesize(self, text: str, references=None, temperature=0.7, top_p=0.8, repetition_penalty=1.1, max_tokens=2048) -> bytes:
        if self._vocoder is None:
            raise RuntimeError("Vocoder not loaded")
        codes = self.generate_codes(text, references, temperature, top_p, repetition_penalty, max_tokens)
        if codes.size(1) == 0:
            raise RuntimeError("No audio generated")
        return self._to_wav_bytes(np.asarray(self._vocoder(codes), dtype=np.float32))

    def synthesize_stream(self, text: str, references=None, chunk_tokens: int = 20, min_first_chunk: int = 10, **kwargs) -> Iterator[bytes]:
        """synthesizer.py:483-584: generation on the caller's thread, vocoding on a worker, two bounded queues."""
        if self._vocoder is None:
            raise RuntimeError("Vocoder not loaded")
        prompt = self._encode_prompt(text, references)
        codes_q: queue.Queue = queue.Queue(maxsize=3)
        audio_q: queue.Queue = queue.Queue(maxsize=3)
        errors: list = []

        def worker():
            try:
                while True:
                    c = codes_q.get()
                    if c is None:
                        break
                    audio = np.asarray(self._vocoder(c), dtype=np.float32)
                    audio_q.put((audio * 32767).astype(np.int16).tobytes())
            except Exception as e:   # re-raised on the caller's thread after the stream drains
                errors.append(e)
            finally:
                audio_q.put(None)

        th = threading.Thread(target=worker, daemon=True)
        th.start()
        try:
            buf, first = [], True
            for col in inference.generate_streaming(model=self._model, prompt=prompt, max_new_tokens=kwargs.get("max_tokens", 2048),
                                                    temperature=kwargs.get("temperature", 0.7), top_p=kwargs.get("top_p", 0.8),
                                                    repetition_penalty=kwargs.get("repetition_penalty", 1.1),
                                                    first_chunk=min_first_chunk, chunk=chunk_tokens):
                col = col.clone()
                col[col < 0] = 0                                  # inference.py:817-818
                buf.append(col)
                if len(buf) >= (min_first_chunk if first else chunk_tokens):
                    codes_q.put(torch.cat(buf, dim=1))
                    buf, first = [], False
                    while not audio_q.empty():
                        a = audio_q.get_nowait()
                        if a is not None:
                            yield a
            if buf:
                codes_q.put(torch.cat(buf, dim=1))
        finally:
            codes_q.put(None)
        th.join()
        while not audio_q.empty():
            a = audio_q.get_nowait()
            if a is not None:
                yield a
        if errors:
            raise errors[0]

    @staticmethod
    def _to_wav_bytes(audio: np.ndarray, sample_rate: int = 44100) -> bytes:
        pcm = (np.clip(audio, -1.0, 1.0) * 32767).astype(np.int16)
        b = io.BytesIO()
        with wave.open(b, "wb") as wf:
            wf.setnchannels(1); wf.setsampwidth(2); wf.setframerate(sample_rate); wf.writeframes(pcm.tobytes())
        return b.getvalue()

    @property
    def sample_rate(self) -> int:
        return 44100

    @property
    def precision(self) -> str:
        return self._precision


def get_instance(model_dir=None, device="cuda", precision="bf16", warmup=True, **kw) -> FishTTS:
    """Process-wide singleton with double-checked locking (synthesizer.py:661-710)."""
    global _instance
    if _instance is not None:
        return _instance
    with _instance_lock:
        if _instance is None:
            _instance = FishTTS(model_dir=model_dir, device=device, precision=precision, warmup=warmup, **kw)
        return _instance


def reset_instance() -> None:
    global _instance
    with _instance_lock:
        _instance = None
