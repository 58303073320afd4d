import sys
from pathlib import Path
import torch
from torch.nn.attention import SDPBackend, sdpa_kernel
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from fish_tts_b200.synthetic import synthetic_prompt
from gpu_common import build_pair, block_noise_source
from helpers import variant_configs
from oracle import dualar_oracle as orc
from test_gpu_parity import oracle_sample

for name in ("s1like", "v15like"):
  for eos in (True, False):
    cfg = variant_configs()[name]
    m, eng, sd = build_pair(cfg, seed=0, bind_kv=False, eos_reachable=eos)
    prompt = synthetic_prompt(cfg, 5, 12, 4, seed=1)
    T, p, rp = 0.7, 0.8, 1.1
    C1 = cfg.num_codebooks + 1
    noise = torch.cat([eng.step_noise(100, s) for s in range(8)])
    eng.set_noise(noise)
    eng.prefill(prompt, 8, T, p, rp)
    toks, fin = eng.collect()
    my_slow, my_fast, nuc = eng.read("slow_logits_raw"), eng.read("fast_logits"), eng.read("nucleus")
    dev = m.device
    t = [torch.tensor(v, device=dev, dtype=torch.float) for v in (T, p, rp)]
    pr = prompt.to(dev); Tlen = pr.size(1)
    m.setup_caches(cfg.max_seq_len)
    tr = []
    blk0 = noise[: eng.noise_per_step]
    with torch.inference_mode(), sdpa_kernel(SDPBackend.MATH):
        for i in range(Tlen - 1):
            orc.forward_generate(m, pr[:, i:i + 1].view(1, C1, 1), torch.tensor([i], device=dev))
        first = orc.decode_one_token_ar(m, pr[:, -1:].view(1, C1, 1), torch.tensor([Tlen - 1], device=dev), *t, None,
                                        noise=block_noise_source(cfg, blk0), stable_ties=True, trace=tr)
    d = (my_slow.float() - tr[0].slow_logits.float().cpu()).abs()
    print(f"{name} eos={eos}: mine first col {toks[:, 0].tolist()} fin={fin} nucleus {nuc.tolist()} | ref {first[:, 0].tolist()} | slow logits max|d| {d.max():.4f} "
          f"| argmax mine {int(my_slow.float().argmax())} ({my_slow.float().max():.3f}) | oracle sampler on my logits -> {oracle_sample(cfg, my_slow, 0, None, T, p, rp, blk0)}")
    lg = tr[0].slow_logits.float().cpu(); top = lg.topk(4)
    print("   ref top4", top.indices.tolist(), [round(v, 3) for v in top.values.tolist()], "noise at top", [round(float(blk0[i]), 4) for i in top.indices.tolist()])
    eng.close()
