"""Where does a decode step's time go?  CPU enqueue cost vs GPU time, slow-only (prefill graph) vs full step."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fish_tts_b200.config import s1_mini_config  # noqa: E402
from fish_tts_b200.engine import DualAREngine  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402

cfg = s1_mini_config()
sd = make_state_dict(cfg, seed=0)
eng = DualAREngine(cfg, sd, device=0, seed=1234)
print("launches (step, prefill position):", eng.launches_per_step())
S = dict(temperature=0.7, top_p=0.8, repetition_penalty=1.1)
for T in (33, 257, 513):
    prompt = synthetic_prompt(cfg, 3, T - 8, 5, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    eng.prefill(prompt, 600, **S)
    t1 = time.perf_counter(); e1.record()
    torch.cuda.synchronize()
    print(f"prefill T={T}: enqueue {1e3 * (t1 - t0):.2f} ms, gpu {e0.elapsed_time(e1):.2f} ms -> {e0.elapsed_time(e1) / T * 1e3:.1f} us/position")
    for n in (64, 256):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); e0.record()
        eng.decode(n)
        t1 = time.perf_counter(); e1.record()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"  decode {n} steps at ctx~{T}: enqueue {1e3 * (t1 - t0):.2f} ms ({1e6 * (t1 - t0) / n:.0f} us/graph), gpu {e0.elapsed_time(e1):.2f} ms "
              f"-> {e0.elapsed_time(e1) / n * 1e3:.1f} us/step, wall {1e3 * (t2 - t0):.2f} ms")
    eng.collect()
