"""Profiling driver (B200_PROFILING.md recipe): s1-mini engine, 223-token prompt, N warm decode steps so the
context is ~600 positions, then exactly `--steps` decode steps inside a cudaProfilerStart/Stop range.

  python tests/profile_step.py --steps 2 && \
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tests/profile_step.py --steps 2
"""
import argparse
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from fish_tts_b200.config import s1_mini_config  # noqa: E402
from fish_tts_b200.engine import DualAREngine  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--context", type=int, default=400, help="decode steps run before the profiled range")
args = ap.parse_args()

cfg = s1_mini_config()
sd = make_state_dict(cfg, seed=0)
eng = DualAREngine(cfg, sd, device=0, seed=1234)
prompt = synthetic_prompt(cfg, 3, 215, 5, seed=1)
eng.prefill(prompt, args.context + args.steps + 8, temperature=0.7, top_p=0.8, repetition_penalty=1.1)
eng.decode(args.context)
torch.cuda.synchronize()
torch.cuda.profiler.start()
t0 = time.perf_counter()
eng.decode(args.steps)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
torch.cuda.profiler.stop()
toks, fin = eng.collect()
print(f"profiled {args.steps} steps at context ~{prompt.size(1) + args.context}: {dt / args.steps * 1e3:.3f} ms/step (wall, includes profiler overhead if any); columns {toks.shape[1]}")
