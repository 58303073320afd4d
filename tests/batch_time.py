"""Timing of the many-column path on the s1-mini shape: tensor-core prefill of a 223-position prompt, and the batched
decode step at several batch sizes (CUDA events; weights 1.4 GB >> L2).  Prints one line per measurement."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fish_tts_b200.config import s1_mini_config  # noqa: E402
from fish_tts_b200.engine import DualAREngine  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402

batches = [int(a) for a in sys.argv[1:]] or [32]
cfg = s1_mini_config()
sd = make_state_dict(cfg, seed=0)
S = dict(temperature=0.7, top_p=0.8, repetition_penalty=1.1)


def ev():
    return torch.cuda.Event(enable_timing=True)


eng = DualAREngine(cfg, sd, device=0, seed=1234)
eng.set_option("prefix_reuse", 0)      # every prefill below pays for the whole prompt
for T in ((223, 512, 1024) if not os.environ.get("BT_SKIP_PREFILL") else ()):
    prompt = synthetic_prompt(cfg, 3, T - 8, 5, seed=1)
    for mode in (0, 1) if T == 223 else (0,):
        eng.set_option("prefill_mode", mode)
        eng.prefill(prompt, 4, **S); torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record(); eng.prefill(prompt, 4, **S); e1.record(); torch.cuda.synchronize()
        print(f"prefill T={T} mode {mode} ({'tcgen05 GEMMs' if mode == 0 else 'one position per launch'}): {e0.elapsed_time(e1):.3f} ms (includes the first decode step)", flush=True)
eng.set_option("prefill_mode", 0)
rng = np.random.default_rng(2)
for B in batches:
    eng2 = DualAREngine(cfg, sd, device=0, seed=1234) if B != batches[0] else eng
    eng2.batch_init(B, 1152)
    lens = rng.integers(64, 513, size=B)
    e0, e1 = ev(), ev()
    e0.record()
    for sl in range(B):
        eng2.batch_prefill(sl, synthetic_prompt(cfg, 3, int(lens[sl]) - 8, 5, seed=10 + sl), 600, **S, seed=100 + sl)
    e1.record(); torch.cuda.synchronize()
    t_pf = e0.elapsed_time(e1)
    eng2.batch_decode(16); torch.cuda.synchronize()
    n = 128
    e0, e1 = ev(), ev()
    e0.record(); eng2.batch_decode(n); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    ctx = float(lens.mean()) + 16 + n / 2
    kv = cfg.weight_bytes()["kv_per_pos"] * (ctx + 1) * B
    by = cfg.weight_bytes()["unique_weights"] + kv
    cols, fin = eng2.batch_collect(0)
    print(f"batched decode B={B}: {ms:.3f} ms/step -> {B / ms * 1e3:.0f} tok/s aggregate; {int(eng2.batch_read('launches')[0])} kernels/step; "
          f"algorithmic {by / 1e9:.2f} GB/step (weights once + KV of {B} requests at mean context {ctx:.0f}) -> {by / ms / 1e6:.0f} GB/s; "
          f"prefill of {B} prompts (mean {lens.mean():.0f} positions) {t_pf:.1f} ms; slot 0 produced {cols.shape[1]} columns", flush=True)
    eng2.close()
