import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from fish_tts_b200.config import s1_mini_config
from fish_tts_b200.synthetic import synthetic_prompt
from gpu_common import TeacherForced, build_pair
cfg = s1_mini_config()
m, eng, sd = build_pair(cfg, seed=0)
tf = TeacherForced(cfg, m, eng, synthetic_prompt(cfg, 5, 12, 4, seed=1), 0.7, 1e-9, 1.0)
sem = slice(cfg.semantic_begin_id, cfg.semantic_end_id + 1)
for s in range(64):
    o = tf.step()
    d = (o["my_slow"].float() - o["ref_slow"].float()).abs()
    rs = o["ref_slow"].float()
    non = torch.cat([rs[:cfg.semantic_begin_id], rs[cfg.semantic_end_id + 1:]])
    if s >= 48 or d.max() > 0.3:
        print(f"step {s}: ref tok {o['ref'][:3].tolist()} mine {o['mine'][:3].tolist()} | max|d| {d.max():.3f} sem max|d| {d[sem].max():.3f} | ref sem max {rs[sem].max():.3f} nonsem max {non.max():.3f} "
              f"| hidden ch0 ref {o['ref_hidden'][0].float():.2f} mine {o['my_hidden'][0].float():.2f} rms {o['ref_hidden'].float().pow(2).mean().sqrt():.2f} hidden max|d| {(o['my_hidden'].float()-o['ref_hidden'].float()).abs().max():.3f}", flush=True)
