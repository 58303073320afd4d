import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from fish_tts_b200.config import s1_mini_config
from fish_tts_b200.synthetic import synthetic_prompt
from gpu_common import TeacherForced, build_pair
cfg = s1_mini_config()
m, eng, sd = build_pair(cfg, seed=0, options={"mega_kernel": 1})
tf = TeacherForced(cfg, m, eng, synthetic_prompt(cfg, 5, 12, 4, seed=1), 0.7, 1e-9, 1.0)
for s in range(14):
    o = tf.step()
    lg = o["my_slow"].float()
    am = int(lg.argmax()); mine = int(o["mine"][0])
    top = torch.topk(lg, 6)
    print(s, "mine", mine, "argmax", am, "n_cand", int(eng.read("n_cand")[0]) if hasattr(eng, "read") else -1, "nucleus", eng.read("nucleus").tolist(),
          "top", [(int(i), round(float(v), 4)) for v, i in zip(top.values, top.indices)], "lg[mine]", float(lg[mine]))
