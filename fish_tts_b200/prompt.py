"""Prompt packing: the numeric half of ``ContentSequence.encode_for_inference`` (fish_tts/models/inference.py:523-640).

The reference builds the (num_codebooks + 1, T) prompt tensor part by part in Python, with one ``.item()`` per VQ code to look its
token id up (inference.py:553-559).  Text -> token ids is the tokenizer's business and stays on the reference side (out of scope);
what is left is arithmetic on id arrays, done here with numpy on whole arrays:

    row 0      token ids; at VQ positions ``codes[0] + semantic_begin_id``   (inference.py:624-627; the per-code lookup table
               ``semantic_id_to_token_id`` is that same affine map, tokenizer.py:84-101)
    rows 1..C  the VQ codes at VQ positions, 0 elsewhere                      (inference.py:628)

``pack_prompt`` packs one request, ``pack_prompts`` a batch of requests of different lengths for the request slots of
``DualAREngine.batch_prefill`` (each slot takes its own length: nothing is padded).
"""
from __future__ import annotations

from typing import Iterable, Sequence, Union

import numpy as np

Part = Union[Sequence[int], np.ndarray]      # 1-D = text token ids, 2-D (num_codebooks, n) = VQ codes


def pack_prompt(parts: Iterable[Part], num_codebooks: int, semantic_begin_id: int, codebook_size: int | None = None) -> np.ndarray:
    """parts in order -> (num_codebooks + 1, T) int32, exactly what ``encode_for_inference`` returns as ``values``."""
    cols = []
    for p in parts:
        a = np.asarray(p.cpu() if hasattr(p, "cpu") else p)
        if a.ndim == 1:                                  # TextPart(tokens=...)
            blk = np.zeros((num_codebooks + 1, a.shape[0]), dtype=np.int32)
            blk[0] = a
        elif a.ndim == 2:                                # VQPart(codes=...)
            if a.shape[0] != num_codebooks:
                raise ValueError(f"VQ part has {a.shape[0]} codebooks, expected {num_codebooks}")
            a = a.astype(np.int32)
            if codebook_size is not None and a.size and (a.min() < 0 or a.max() >= codebook_size):
                raise ValueError("VQ code out of range")
            blk = np.empty((num_codebooks + 1, a.shape[1]), dtype=np.int32)
            blk[0] = a[0] + semantic_begin_id
            blk[1:] = a
        else:
            raise ValueError(f"Unsupported part with {a.ndim} dimensions")
        cols.append(blk)
    if not cols:
        return np.zeros((num_codebooks + 1, 0), dtype=np.int32)
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def pack_prompts(requests: Iterable[Iterable[Part]], num_codebooks: int, semantic_begin_id: int, codebook_size: int | None = None):
    """A batch of requests -> (list of (num_codebooks + 1, T_i) int32 prompts, int32 array of the T_i)."""
    out = [pack_prompt(r, num_codebooks, semantic_begin_id, codebook_size) for r in requests]
    return out, np.asarray([p.shape[1] for p in out], dtype=np.int32)
