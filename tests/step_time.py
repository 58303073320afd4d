"""Decode step time of the s1-mini engine under a list of environment settings, one process (weights generated once).

  python tests/step_time.py "" "DUALAR_L2_WINDOW=1" "DUALAR_L2_WINDOW=1 DUALAR_L2_HIT=0.6"

Each argument is a space-separated list of VAR=value pairs applied before the engine is built (the switches are read at
dualar_finalize).  Prints ms/step over `--steps` decode steps after a 223-position prompt (CUDA events, weights >> L2).
"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fish_tts_b200.config import s1_mini_config  # noqa: E402
from fish_tts_b200.engine import DualAREngine  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
steps = 512
for a in sys.argv[1:]:
    if a.startswith("--steps="):
        steps = int(a.split("=")[1])
cfg = s1_mini_config()
sd = make_state_dict(cfg, seed=0)
prompt = synthetic_prompt(cfg, 3, 215, 5, seed=1)
for setting in (args or [""]):
    applied = {}
    for kv in setting.split():
        k, v = kv.split("=", 1)
        applied[k] = os.environ.get(k)
        os.environ[k] = v
    eng = DualAREngine(cfg, sd, device=0, seed=1234)
    eng.prefill(prompt, steps + 80, temperature=0.7, top_p=0.8, repetition_penalty=1.1)
    eng.decode(64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.decode(steps)
    e1.record()
    torch.cuda.synchronize()
    toks, fin = eng.collect()
    print(f"[{setting or 'default'}] {e0.elapsed_time(e1) / steps:.4f} ms/step over {steps} steps ({toks.shape[1]} columns)", flush=True)
    eng.close()
    for k, v in applied.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
