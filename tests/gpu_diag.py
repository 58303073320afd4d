"""Diagnostic run for the GPU box (not a pytest): prints per-stage numbers so a failing kernel can be
located from one gpurun call.  python tests/gpu_diag.py [tiny|s1] > gpurun_out/diag.log"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from fish_tts_b200.config import s1_mini_config, tiny_config  # noqa: E402
from fish_tts_b200.synthetic import synthetic_prompt  # noqa: E402
from gpu_common import TeacherForced, build_pair  # noqa: E402
from helpers import bf16_ulp, variant_configs  # noqa: E402


def cmp(name, a, b):
    a, b = a.float().cpu().flatten(), b.float().cpu().flatten()
    d = (a - b).abs()
    ulp = (d / bf16_ulp(b)).max().item()
    print(f"   {name:12s} n={a.numel():7d} max|d|={d.max().item():.5f} max_ulp={ulp:.1f} mismatched={(d > 0).float().mean().item():.4f} "
          f"ref_absmax={b.abs().max().item():.3f} nan={int(torch.isnan(a).sum())}", flush=True)


def run(cfg, label, n_steps, modes, with_alt=False):
    print(f"=== {label}", flush=True)
    t0 = time.time()
    m, eng, sd = build_pair(cfg)
    alt = None
    if with_alt:
        from oracle import dualar_oracle as orc
        alt = orc.OracleModel.build(cfg, sd, device="cpu")
    print(f"   built in {time.time() - t0:.1f}s; launches/step {eng.launches_per_step()} weight bytes {eng.weight_bytes()}", flush=True)
    prompt = synthetic_prompt(cfg, 5, 12, 4, seed=1)
    for (T, p, rp) in modes:
        tf = TeacherForced(cfg, m, eng, prompt, T, p, rp, alt=alt)
        agree = 0
        agree_alt = 0
        for s in range(n_steps):
            o = tf.step()
            same = torch.equal(o["mine"], o["ref"])
            agree += same
            if alt is not None:
                agree_alt += torch.equal(o["alt"], o["ref"])
                sem = slice(cfg.semantic_begin_id, cfg.semantic_end_id + 1)
                dm = (o["my_slow"].float() - o["ref_slow"].float()).abs()
                da = (o["alt_slow"].float() - o["ref_slow"].float()).abs()
                top2 = o["ref_slow"].float()[sem].topk(2).values
                print(f"  step {s}: |mine-cuda| sem max {dm[sem].max():.4f} mean {dm[sem].mean():.5f} all max {dm.max():.4f} | |cpu-cuda| sem max {da[sem].max():.4f} "
                      f"mean {da[sem].mean():.5f} all max {da.max():.4f} | hidden mine {(o['my_hidden'].float()-o['ref_hidden'].float()).abs().max():.3f} "
                      f"cpu {(o['alt_hidden'].float()-o['ref_hidden'].float()).abs().max():.3f} | top2 gap {top2[0]-top2[1]:.4f} | tok mine==cuda {same} cpu==cuda {torch.equal(o['alt'], o['ref'])}", flush=True)
            if s < 3 or not same:
                print(f"  step {s} T={T} p={p} rp={rp} tokens {'==' if same else '!='} mine {o['mine'].tolist()} ref {o['ref'].tolist()} nucleus {o['nucleus'].tolist()}")
                cmp("hidden", o["my_hidden"], o["ref_hidden"])
                cmp("slow_logits", o["my_slow"], o["ref_slow"])
                k = 0
                while k < cfg.num_codebooks - 1 and o["mine"][k + 1] == o["ref"][k + 1]:
                    k += 1
                cmp("fast_logits", o["my_fast"][: max(k, 1)], o["ref_fast"][: max(k, 1)])
        print(f"  T={T} p={p} rp={rp}: {agree}/{n_steps} steps token-identical (cpu oracle vs cuda oracle: {agree_alt}/{n_steps})", flush=True)
    # one-layer intermediates
    eng.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    torch.manual_seed(0)
    modes = [(0.7, 1e-9, 1.0), (0.7, 0.8, 1.1)]
    if what == "tiny":
        one = tiny_config(n_layer=1, n_fast_layer=1)
        m, eng, sd = build_pair(one)
        prompt = synthetic_prompt(one, 5, 12, 4, seed=1)
        tf = TeacherForced(one, m, eng, prompt, 0.7, 1e-9, 1.0)
        o = tf.step()
        print("=== one-layer tiny: intermediates of the single slow layer (engine buffers vs oracle recomputation)")
        # recompute the oracle's intermediates for the same step
        import torch.nn.functional as F
        from oracle import dualar_oracle as orc
        cmp("hidden", o["my_hidden"], o["ref_hidden"])
        cmp("slow_logits", o["my_slow"], o["ref_slow"])
        cmp("fast_logits", o["my_fast"], o["ref_fast"])
        print("   tokens mine", o["mine"].tolist(), "ref", o["ref"].tolist())
        eng.close()
        for name, cfg in variant_configs().items():
            run(cfg, f"tiny/{name}", 24, modes)
    else:
        run(s1_mini_config(), "s1-mini", 10, modes, with_alt=True)
