"""Persistent batched step (bstep.cuh) against the per-kernel graph: identical bits, then timing.  Runs on the B200 box.
    python tests/bstep_check.py [tiny|s1mini] [batch] [steps]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fish_tts_b200.config import s1_mini_config, tiny_config  # noqa: E402
from fish_tts_b200.engine import DualAREngine  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
cfg = tiny_config() if model == "tiny" else s1_mini_config()
sd = make_state_dict(cfg, seed=0)
eng = DualAREngine(cfg, sd, device=0, seed=1234)
eng.set_option("prefix_reuse", 0)
Sb = 256 if model == "tiny" else 1152
eng.batch_init(B, Sb)
rng = np.random.default_rng(3)
lens = rng.integers(12, 60 if model == "tiny" else 400, size=B)
prompts = [synthetic_prompt(cfg, 3, int(lens[i]) - 8, 5, seed=10 + i) for i in range(B)]


def run(persist):
    eng.set_option("batch_persistent", persist)
    for sl in range(B):
        eng.batch_prefill(sl, prompts[sl], n_steps + 4, 0.7, 0.8, 1.1, seed=100 + sl)
    out = []
    for s in range(n_steps):
        eng.batch_decode(1)
        out.append((eng.batch_read("tokens").clone(), eng.batch_read("slow_logits_raw").clone(), eng.batch_read("fast_logits").clone(), eng.batch_read("hidden").clone()))
    for sl in range(B):
        eng.batch_collect(sl)
        eng.batch_release(sl)
    return out


a = run(0)
print("per-kernel path done;", int(eng.batch_read("launches")[0]), "launches", flush=True)
b = run(1)
print("persistent path done;", int(eng.batch_read("launches")[0]), "launches", flush=True)
bad = 0
for s in range(n_steps):
    for name, x, y in zip(("tokens", "slow_logits", "fast_logits", "hidden"), a[s], b[s]):
        if not torch.equal(x, y):
            bad += 1
            d = (x.float() - y.float()).abs()
            print(f"step {s} {name}: {int((x != y).sum())} of {x.numel()} differ, max |d| {float(d.max()):.4g}", flush=True)
    if bad > 6:
        break
print("IDENTICAL" if bad == 0 else f"MISMATCH ({bad})", flush=True)

# timing
for persist in (0, 1):
    eng.set_option("batch_persistent", persist)
    for sl in range(B):
        eng.batch_prefill(sl, prompts[sl], 400, 0.7, 0.8, 1.1, seed=100 + sl)
    eng.batch_decode(8); torch.cuda.synchronize()
    n = 64 if model != "tiny" else 32
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.batch_decode(n); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{model} B={B} persistent={persist}: {ms:.3f} ms/step -> {B / ms * 1e3:.0f} tok/s", flush=True)
    if persist and os.environ.get("DUALAR_BS_TIMELINE") == "1":
        kinds = eng.batch_read("bstep_kinds").numpy(); tl = eng.batch_read("bstep_timeline").numpy().astype(np.float64)
        split = int(np.nonzero((kinds & 255) == 6)[0][np.argmax((kinds >> 8)[(kinds & 255) == 6])]) + 1      # the LM head is the GEMM with most units
        names = ["embed", "norm", "qkv_post", "attn", "fast_attn", "fast_sample", "gemm"]
        work = tl[:, 1] - tl[:, 0]
        nxt = np.roll(tl[:, 0], -1) - tl[:, 1]      # barrier: end of own work -> start of the next phase
        nxt[split - 1] = 0; nxt[-1] = 0
        clk = 1.9e3      # cycles per us (approx.)
        print(f"timeline of CTA 0 (cycles / {clk:.0f} = us): launch 1 = phases [0, {split}), launch 2 = [{split}, {len(kinds)})")
        for kd in range(7):
            m = (kinds & 255) == kd
            if m.any():
                print(f"  {names[kd]:12s} n {int(m.sum()):4d}  own work mean {work[m].mean() / clk:7.2f} us (max {work[m].max() / clk:7.2f})   barrier after it mean {nxt[m].mean() / clk:6.2f} us   total {(work[m].sum() + nxt[m].sum()) / clk / 1e3:7.3f} ms")
        gm = (kinds & 255) == 6
        for units in sorted(set((kinds >> 8)[gm])):
            m = gm & ((kinds >> 8) == units)
            print(f"    gemm with {units:5d} units: n {int(m.sum()):4d} own work {work[m].mean() / clk:7.2f} us, barrier {nxt[m].mean() / clk:6.2f} us")
        print(f"  launch 1 {(tl[split - 1, 1] - tl[0, 0]) / clk / 1e3:.3f} ms, launch 2 {(tl[-1, 1] - tl[split, 0]) / clk / 1e3:.3f} ms")
    for sl in range(B):
        eng.batch_collect(sl)
        eng.batch_release(sl)
eng.close()
