import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from fish_tts_b200.synthetic import synthetic_prompt
from gpu_common import TeacherForced, build_pair
from helpers import variant_configs
cfg = variant_configs()["biased"]
for flag in (1, 0):
    m, eng, sd = build_pair(cfg, seed=0, options={"mega_kernel": flag})
    tf = TeacherForced(cfg, m, eng, synthetic_prompt(cfg, 5, 12, 4, seed=1), 0.7, 1e-9, 1.0)
    o = tf.step()
    lg = o["my_slow"].float()
    mx = lg.max()
    ties = (lg == mx).nonzero().flatten().tolist()
    print("mega" if flag else "phase", "mine", o["mine"].tolist(), "ref", o["ref"].tolist(), "max", mx.item(), "argmax ties", ties[:10], "nucleus", o["nucleus"].tolist())
    print("  logit[mine]", lg[int(o["mine"][0])].item(), "ref logits max", o["ref_slow"].float().max().item(), "maxdiff", (lg - o["ref_slow"].float()).abs().max().item())
    eng.close()
m, eng, sd = build_pair(cfg, seed=0, options={"mega_kernel": 1})
tf = TeacherForced(cfg, m, eng, synthetic_prompt(cfg, 5, 12, 4, seed=1), 0.7, 1e-9, 1.0)
o = tf.step()
cand = eng.read("cand")
lgb = o["my_slow"].view(torch.int16).to(torch.int32) & 0xFFFF
n = int((cand != 0).sum())
idx = ((cand >> 30) & 0x3FFFF)[:n]
inv = ((cand >> 48) & 0xFFFF)[:n]
print("n", n, "idx ascending:", bool((idx[1:] > idx[:-1]).all()), "idx first/last", idx[:8].tolist(), idx[-4:].tolist())
key = torch.where((lgb & 0x8000) != 0, (~lgb) & 0xFFFF, lgb | 0x8000)
exp_inv = 0xFFFF - key[idx.long()]
print("keys match:", bool((exp_inv == inv).all()), "tags", set((cand[:n] & 0x3FFFFFFF).tolist()))
best = int(inv.argmin()); print("min inv at slot", best, "idx", int(idx[best]))
