"""DUALAR_TIMELINE=1 python tests/timeline.py : in-kernel %globaltimer stamps of one decode step (block 0 of each kernel)."""
import os
import sys
from pathlib import Path

import torch

os.environ["DUALAR_TIMELINE"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fish_tts_b200.config import s1_mini_config  # noqa: E402
from fish_tts_b200.engine import DualAREngine  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402

cfg = s1_mini_config()
eng = DualAREngine(cfg, make_state_dict(cfg, seed=0), device=0, seed=1234)
n_step, n_pre = eng.launches_per_step()
prompt = synthetic_prompt(cfg, 3, 215, 5, seed=1)
eng.prefill(prompt, 600, temperature=0.7, top_p=0.8, repetition_penalty=1.1)
eng.decode(300)
torch.cuda.synchronize()
tl = eng.read("timeline")[:n_step].numpy()
g = tl[:, :4].astype("int64")
t0 = g[0, 0]
names = []
L = cfg.n_layer
names.append("embed")
for i in range(L):
    names += [f"L{i}.qkv", f"L{i}.attn", f"L{i}.wo", f"L{i}.w13", f"L{i}.w2"]
names += ["head", "select"]
for p in range(cfg.num_codebooks):
    for l in range(cfg.n_fast_layer):
        names += [f"F{p}.{l}.qkv", f"F{p}.{l}.wo", f"F{p}.{l}.w13", f"F{p}.{l}.w2"]
    if p:
        names.append(f"F{p}.head")
assert len(names) == n_step, (len(names), n_step)
print("slot name           entry    wait_ret  pro_done  end(b0)   | gap_from_prev_end  wait-entry  pro-wait  end-pro   (us, relative to step start)")
prev_end = None
import collections
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])
for i, nm in enumerate(names):
    e, w, p, x = [(v - t0) / 1e3 for v in g[i]]
    period = (g[i + 1, 0] - g[i, 0]) / 1e3 if i + 1 < n_step else float("nan")
    kind = nm.split(".")[-1] if "." in nm else nm
    kind = ("F." if nm.startswith("F") else "S.") + kind
    a = agg[kind]; a[0] += 1; a[1] += w - e; a[2] += p - w; a[3] += x - p; a[4] += (x - e)
    if i + 1 < n_step:
        a[5] += (g[i + 1, 1] - g[i, 1]) / 1e3     # wait-return to next wait-return = the serial period
    if i < 14 or (145 <= i < 170) or i > n_step - 20:
        print(f"{i:4d} {nm:12s} {e:9.2f} {w:9.2f} {p:9.2f} {x:9.2f}   | {'' if prev_end is None else f'{e - prev_end:8.2f}':>8s} {w - e:10.2f} {p - w:9.2f} {x - p:8.2f}")
    prev_end = x
print(f"step total (first entry -> last end): {(g[-1, 3] - t0) / 1e3:.1f} us")
print("kind        n   wait-entry  pro-wait  end-pro  total(b0)  serial period (wait_ret -> next wait_ret)")
for k, a in agg.items():
    n = a[0]
    print(f"{k:10s} {n:3d} {a[1] / n:9.2f} {a[2] / n:9.2f} {a[3] / n:8.2f} {a[4] / n:9.2f} {a[5] / n:9.2f}   sum {a[5]:8.1f}")
