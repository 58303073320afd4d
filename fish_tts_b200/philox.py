"""Philox4x32-10 in numpy -- the host-side twin of ``exp1_noise`` in csrc/common.cuh.

The sampler's Exp(1) draws (``multinomial_sample_one_no_sync``, fish_tts/models/inference.py:24-27) come
from a counter-based stream keyed by (seed, step, head, element), so a host, the oracle and the kernels
can all reproduce them: counter = (element, head, step, 0), key = (seed_lo, seed_hi),
u = ((x0 >> 9) + 0.5) * 2^-23 (exact in fp32, strictly inside (0,1), so q > 0), q = bf16(-log(u)).
"""

from __future__ import annotations

import numpy as np
import torch

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _MASK, p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniform23(seed: int, step: int, head: int, n: int) -> np.ndarray:
    elem = np.arange(n, dtype=np.uint64)
    z = np.zeros(n, dtype=np.uint64)
    x0, _, _, _ = philox4x32_10(elem, z + np.uint64(head), z + np.uint64(step), z, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return ((x0 >> np.uint64(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)


def exp1_noise(seed: int, step: int, head: int, n: int) -> torch.Tensor:
    """bf16 Exp(1) draws for one head of one step (what dualar_fill_noise produces on the device)."""
    u = uniform23(seed, step, head, n)
    q = (-np.log(u.astype(np.float64))).astype(np.float32)
    return torch.from_numpy(q).to(torch.bfloat16)


def step_noise(cfg, seed: int, step: int) -> torch.Tensor:
    """One step's noise block: slow head (vocab_size) then fast heads 1..C-1 (min(1024, codebook_size) each)."""
    fv = min(1024, cfg.codebook_size)
    parts = [exp1_noise(seed, step, 0, cfg.vocab_size)]
    parts += [exp1_noise(seed, step, k, fv) for k in range(1, cfg.num_codebooks)]
    return torch.cat(parts)


def oracle_noise_fn(cfg, seed: int):
    """call index -> (step, head) in the order the reference samples: slow head, then codebooks 1..C-1."""
    C = cfg.num_codebooks

    def fn(call: int, n: int) -> torch.Tensor:
        return exp1_noise(seed, call // C, call % C, n)

    return fn
