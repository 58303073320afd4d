"""ORACLE SUPPORT (test infrastructure) -- a numpy model of the CUDA sampler's *definition*
(fish_tts_b200/csrc/sampler.cuh), so the design can be checked against the reference's sampler
(inference.py:30-80 as restated in dualar_oracle.sample) on the CPU, without a GPU:
  * ties ordered by ascending index,
  * nucleus from an exact fixed-point (2^-44) running sum of the bf16 probabilities,
  * everything else with the reference's rounding points.
Used by tests/test_sampler_model.py only.
"""

from __future__ import annotations

import numpy as np
import torch

FIX = float(2 ** 44)


def _bf(x: np.ndarray) -> np.ndarray:
    """round float32 array to the bf16 grid (RNE), returned as float32"""
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).bfloat16().float().numpy()


def cmax_from_top_p(top_p: float) -> int:
    t = torch.tensor(top_p, dtype=torch.float32).bfloat16()
    tb = int(t.view(torch.int16).item()) & 0xFFFF
    nxt = torch.tensor([tb + 1], dtype=torch.int32).to(torch.int16).view(torch.bfloat16).float().item()
    mid = np.float32(0.5) * np.float32(t.float().item()) + np.float32(0.5) * np.float32(nxt)
    M = int(np.float32(mid) * np.float32(FIX))
    return (M - 1 if M else 0) if (tb & 1) else M


def sample(logits_bf16: torch.Tensor, temperature: float, top_p: float, rep_penalty: float,
           prev_ids, noise_bf16: torch.Tensor, cpu_semantics: bool = True):
    """-> (token index, nucleus size).  logits/noise are 1-D bf16 tensors.
    cpu_semantics: temperature / repetition penalty stay fp32 (torch CPU) instead of being cast to bf16 (torch CUDA)."""
    z = logits_bf16.float().numpy().copy()
    if prev_ids is not None:
        rp = np.float32(rep_penalty) if cpu_semantics else _bf(np.float32(rep_penalty))
        ids = np.unique(np.asarray(prev_ids, dtype=np.int64))
        ids = ids[(ids >= 0) & (ids < z.size)]
        s = z[ids]
        z[ids] = np.where(s < 0, _bf(s * rp), _bf(s / rp))
    m = z.max()
    e = np.exp((z - m).astype(np.float32)).astype(np.float32)
    S = np.float32(e.sum(dtype=np.float64))
    p = _bf(e / S)
    w = (p.astype(np.float64) * FIX).astype(np.int64)            # exact for bf16 values >= 2^-36
    order = np.lexsort((np.arange(z.size), -z))                   # logit desc, index asc
    cum = np.cumsum(w[order])
    n_keep = max(1, int((cum <= cmax_from_top_p(top_p)).sum()))
    kept = order[:n_keep]
    T = np.float32(max(temperature, 1e-5)) if cpu_semantics else _bf(np.float32(max(temperature, 1e-5)))
    z2 = _bf(z[kept] / T)
    e2 = np.exp((z2 - z2[0]).astype(np.float32)).astype(np.float32)
    p2 = _bf(e2 / np.float32(e2.sum(dtype=np.float64)))
    r = _bf(p2 / noise_bf16.float().numpy()[kept])
    best_r, best_i = np.float32(0.0), 0                          # removed tokens: r = 0, ties -> index 0
    rmax = r.max()
    if rmax > 0:
        cand = kept[r == rmax]
        best_i = int(cand.min())
        best_r = rmax
    return best_i, n_keep


def nucleus_sorted(z: np.ndarray, w: np.ndarray, c_max: int) -> np.ndarray:
    """kept mask by the definition: prefix of the (logit desc, index asc) order whose inclusive fixed-point cumulative
    weight stays <= c_max; the first item is always kept."""
    order = np.lexsort((np.arange(z.size), -z))
    cum = np.cumsum(w[order])
    n_keep = max(1, int((cum <= c_max).sum()))
    kept = np.zeros(z.size, dtype=bool)
    kept[order[:n_keep]] = True
    return kept


def nucleus_binned(z: np.ndarray, w: np.ndarray, c_max: int, n_bins: int = 512, per_unit: int = 64) -> np.ndarray:
    """kept mask the way sample_binned() (fish_tts_b200/csrc/sampler.cuh) finds it, without sorting: items binned by
    their distance below the maximum (monotone in the logit), exact per-bin weight sums, one scan for the bin in
    which the cumulative weight crosses c_max; bins ahead are kept whole, bins behind dropped whole, and only the
    items of the cut bin are ranked against each other."""
    m = np.float32(z.max())
    b = np.minimum(np.maximum((m - z.astype(np.float32)) * np.float32(per_unit), np.float32(0)), np.float32(n_bins - 1)).astype(np.int64)
    hist = np.zeros(n_bins, dtype=np.int64)
    np.add.at(hist, b, w)
    cnt = np.bincount(b, minlength=n_bins)
    incl = np.cumsum(hist)
    excl = incl - hist
    cross = np.nonzero((excl <= c_max) & (incl > c_max))[0]
    kept = np.zeros(z.size, dtype=bool)
    if cross.size == 0:                      # the whole mass fits: everything is kept
        kept[:] = True
        return kept
    cut = int(cross[0])
    kept[b < cut] = True
    before, cnt_before = int(excl[cut]), int(cnt[:cut].sum())
    idx = np.nonzero(b == cut)[0]
    for i in idx:                            # composite order inside the cut bin: logit desc, index asc
        ahead = idx[(-z[idx] < -z[i]) | ((z[idx] == z[i]) & (idx < i))]
        g = before + int(w[ahead].sum())
        if g + int(w[i]) <= c_max or (ahead.size == 0 and cnt_before == 0):
            kept[i] = True
    return kept
