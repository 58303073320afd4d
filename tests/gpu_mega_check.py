"""GPU bring-up check: persistent whole-step kernel vs the per-phase kernels, bit for bit, plus a quick timing."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from fish_tts_b200.config import fish_speech_1_5_config, s1_mini_config  # noqa: E402
from fish_tts_b200.engine import DualAREngine  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402
from helpers import variant_configs  # noqa: E402

names = sys.argv[1:] or ["s1like", "v15like", "biased", "s1mini"]
for name in names:
    cfg = s1_mini_config() if name == "s1mini" else fish_speech_1_5_config() if name == "v15" else variant_configs()[name]
    sd = make_state_dict(cfg, seed=0)
    prompt = synthetic_prompt(cfg, 5, 12, 4, seed=1)
    outs = []
    for flag in (1, 0):
        eng = DualAREngine(cfg, sd, device=0, seed=5, options={"mega_kernel": flag})
        t0 = time.time()
        try:
            toks = eng.generate(prompt, int(__import__('os').environ.get('NTOK', '24')), 0.7, 0.8, 1.1)
        except Exception as ex:
            print(name, "mega" if flag else "phase", "FAILED:", ex)
            outs.append(None)
            eng.close()
            continue
        dt = time.time() - t0
        outs.append((toks, eng.read("fast_logits").clone(), eng.read("slow_logits_raw").clone(), eng.read("hidden").clone(), eng.launches_per_step()))
        print(name, "mega" if flag else "phase", "launches/step", eng.launches_per_step(), f"generate 24: {dt*1e3:.1f} ms")
        eng.close()
    if outs[0] is None or outs[1] is None:
        continue
    a, b = outs
    print(name, "tokens equal:", bool((a[0] == b[0]).all()), "fast_logits equal:", torch.equal(a[1], b[1]),
          "slow_logits equal:", torch.equal(a[2], b[2]), "hidden equal:", torch.equal(a[3], b[3]))
    if not (a[0] == b[0]).all():
        import numpy as np
        bad = np.argwhere(np.asarray(a[0]) != np.asarray(b[0]))
        print("  first mismatches (row, col):", bad[:8].tolist())
        print("  mega :", np.asarray(a[0])[:, :6].tolist())
        print("  phase:", np.asarray(b[0])[:, :6].tolist())
        sl_a, sl_b = a[2].float(), b[2].float()
        ta, tb = int(np.asarray(a[0])[0, 0]), int(np.asarray(b[0])[0, 0])
        print(f"  slow logit of mega's token: mega {sl_a[ta]:.4f} phase {sl_b[ta]:.4f}; of phase's token: mega {sl_a[tb]:.4f} phase {sl_b[tb]:.4f}")
    if not torch.equal(a[2], b[2]):
        d = (a[2].float() - b[2].float()).abs()
        print("  slow logits max diff", d.max().item(), "n diff", int((d > 0).sum()))
    d = (a[2].float() - b[2].float())
    big = b[2].float().abs() > 16
    if big.any():
        print(f"  slow logits mega - phase: all mean {d.mean():+.4f} max|d| {d.abs().max():.4f}; |x|>16: n={int(big.sum())} mean {d[big].mean():+.4f} std {d[big].std():.4f}; |x|<=16: mean {d[~big].mean():+.4f} max|d| {d[~big].abs().max():.4f}")
    dh = (a[3].float() - b[3].float())
    print(f"  hidden mega - phase: max|d| {dh.abs().max():.4f} n diff {int((dh != 0).sum())}/{dh.numel()}")
