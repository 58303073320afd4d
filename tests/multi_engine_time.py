"""Do two (or more) batched decode chains overlap on one B200?  E engines x B slots, each engine on its own stream.
    python tests/multi_engine_time.py E B [steps]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fish_tts_b200.config import s1_mini_config  # noqa: E402
from fish_tts_b200.engine import DualAREngine  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n = int(sys.argv[3]) if len(sys.argv) > 3 else 96
cfg = s1_mini_config()
sd = make_state_dict(cfg, seed=0)
rng = np.random.default_rng(2)
engs, streams = [], []
for i in range(E):
    eng = DualAREngine(cfg, sd, device=0, seed=1234 + i)
    eng.set_option("prefix_reuse", 0)
    eng.batch_init(B, 1152)
    lens = rng.integers(64, 513, size=B)
    for sl in range(B):
        eng.batch_prefill(sl, synthetic_prompt(cfg, 3, int(lens[sl]) - 8, 5, seed=10 + sl + 100 * i), 600, 0.7, 0.8, 1.1, seed=100 + sl)
    engs.append(eng); streams.append(torch.cuda.Stream())
torch.cuda.synchronize()
for k in (1, E):
    for eng, st in zip(engs[:k], streams[:k]):
        with torch.cuda.stream(st):
            eng.batch_decode(8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in streams[:k]:
        st.wait_event(e0)
    for i in range(n):      # interleave the graph launches so neither stream's queue runs dry
        for eng, st in zip(engs[:k], streams[:k]):
            with torch.cuda.stream(st):
                eng.batch_decode(1)
    for st in streams[:k]:
        torch.cuda.current_stream().wait_stream(st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{k} engine(s) x {B} slots: {ms:.3f} ms per round of steps -> {k * B / ms * 1e3:.0f} tok/s aggregate", flush=True)
for eng in engs:
    eng.close()
