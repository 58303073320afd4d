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
if n_step == len(names) + 1:
    names.append("fast_ar")
else:
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
    kind = ("F." if nm.startswith("F") and nm != "fast_ar" else "S.") + kind
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

if names[-1] == "fast_ar":
    full = eng.read("timeline").numpy()
    nph = cfg.num_codebooks * cfg.n_fast_layer * 4 + cfg.num_codebooks - 1
    ph = full[192:192 + nph, :3].astype("int64")
    base = ph[0, 0]
    kinds = []
    for p in range(cfg.num_codebooks):
        for l in range(cfg.n_fast_layer):
            kinds += ["qkv", "wo", "w13", "w2"]
        if p:
            kinds.append("head")
    print("fast_ar phases (block 0): start -> staged -> pairs done, us")
    import collections
    ag = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i in range(nph):
        st_, sg, dn = [(v - base) / 1e3 for v in ph[i]]
        nxt = (ph[i + 1, 0] - base) / 1e3 if i + 1 < nph else dn
        a_ = ag[kinds[i]]; a_[0] += 1; a_[1] += sg - st_; a_[2] += dn - sg; a_[3] += nxt - st_
        if i < 20 or i > nph - 8:
            print(f"  ph {i:3d} {kinds[i]:5s} start {st_:9.2f} staged +{sg - st_:6.2f} pairs +{dn - sg:6.2f} period {nxt - st_:6.2f}")
    for k, a_ in ag.items():
        print(f"  {k:5s} n={a_[0]:3d} stage {a_[1] / a_[0]:6.2f} pairs {a_[2] / a_[0]:6.2f} period {a_[3] / a_[0]:6.2f}  sum {a_[3]:8.1f}")
