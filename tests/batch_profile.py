"""Profiling driver for the batched decode step (B200_PROFILING.md recipe): s1-mini, B slots with mixed prompts, a few warm
steps, then `--steps` batched steps inside a cudaProfilerStart/Stop range.

  python tests/batch_profile.py --batch 32 --steps 1 && \
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tests/batch_profile.py --batch 32 --steps 1
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fish_tts_b200.config import s1_mini_config  # noqa: E402
from fish_tts_b200.engine import DualAREngine  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--prompt", type=int, default=0, help="fixed prompt length (0: uniform in [64, 512], seed 2)")
args = ap.parse_args()
cfg = s1_mini_config()
sd = make_state_dict(cfg, seed=0)
eng = DualAREngine(cfg, sd, device=0, seed=1234)
eng.batch_init(args.batch, 1152)
rng = np.random.default_rng(2)
lens = rng.integers(64, 513, size=args.batch) if not args.prompt else np.full(args.batch, args.prompt)
for sl in range(args.batch):
    eng.batch_prefill(sl, synthetic_prompt(cfg, 3, int(lens[sl]) - 8, 5, seed=10 + sl), 600, 0.7, 0.8, 1.1, seed=100 + sl)
eng.batch_decode(8)
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.batch_decode(args.steps); e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"batched decode B={args.batch}: {e0.elapsed_time(e1) / args.steps:.3f} ms/step over {args.steps} step(s), mean prompt {lens.mean():.0f}")
