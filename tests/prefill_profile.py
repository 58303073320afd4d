"""Profiling driver for the tensor-core prefill (B200_PROFILING.md recipe): s1-mini, a T-position prompt prefilled once to warm up,
then once more inside a cudaProfilerStart/Stop range.

  python tests/prefill_profile.py --positions 1024 && \
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 8 -c 4 \
      -o gpurun_out/prof_gemm_tc python tests/prefill_profile.py --positions 1024
"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fish_tts_b200.config import s1_mini_config  # noqa: E402
from fish_tts_b200.engine import DualAREngine  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--positions", type=int, default=1024)
args = ap.parse_args()
cfg = s1_mini_config()
eng = DualAREngine(cfg, make_state_dict(cfg, seed=0), device=0, seed=1234)
eng.set_option("prefix_reuse", 0)
prompt = synthetic_prompt(cfg, 3, args.positions - 8, 5, seed=1)
eng.prefill(prompt, 2, 0.7, 0.8, 1.1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.prefill(prompt, 2, 0.7, 0.8, 1.1); e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"prefill of {args.positions} positions: {e0.elapsed_time(e1):.3f} ms, {int(eng.read('prefill_launches')[0])} kernels")
