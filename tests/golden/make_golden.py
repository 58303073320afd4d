"""Generate the golden fixtures of the decode path by running the UNMODIFIED reference
(/root/reference, CPU, eager, bf16) on the seeded checkpoints of fish_tts_b200.synthetic.

    python tests/golden/make_golden.py            # writes tests/golden/*.pt

Run here in the build container only (the GPU box has no /root/reference).  The fixtures pin
``oracle/dualar_oracle.py`` (tests/test_oracle_golden.py) and are a second parity target for the CUDA
path (tests/test_gpu_parity.py).  Noise for ``multinomial_sample_one_no_sync`` is the explicit Philox
stream of fish_tts_b200.philox, injected by patching that module-level function of the reference.
"""
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from fish_tts_b200 import philox  # noqa: E402
from fish_tts_b200.config import s1_mini_config  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402
from helpers import variant_configs  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = Path(__file__).resolve().parent
SAMPLING = {"sampled": (0.7, 0.8, 1.1), "greedy": (0.7, 1e-9, 1.0), "hot": (1.0, 1.0, 1.5)}


def run_case(cfg, sd, prompt, n_new, mode, seed, full_logits=True):
    llama, inf = rh.import_reference()
    with tempfile.TemporaryDirectory() as d:
        model, _ = rh.load_reference_model(cfg, sd, d)
    T, p, rp = SAMPLING[mode]
    with rh.RecordingStep(model, inf, philox.oracle_noise_fn(cfg, seed)) as rec:
        y = inf.generate(model=model, prompt=prompt.clone(), max_new_tokens=n_new, audio_masks=None, audio_parts=None,
                         decode_one_token=rec.step, temperature=T, top_p=p, repetition_penalty=rp)
    steps = rec.steps
    out = {
        "mode": mode, "temperature": T, "top_p": p, "repetition_penalty": rp, "noise_seed": seed,
        "prompt": prompt.clone(), "seq": y.clone(),
        "tokens": torch.stack([s["tokens"] for s in steps]),             # (n_steps, C+1)
        "hidden": torch.stack([s["hidden"] for s in steps]),             # (n_steps, dim)
        "fast_logits": torch.stack([torch.stack(s["fast_logits"]) for s in steps]),   # (n_steps, C-1, fv)
    }
    sl = torch.stack([s["slow_logits"] for s in steps])                  # (n_steps, V) raw (before penalty)
    if full_logits:
        out["slow_logits"] = sl
    else:   # s1-mini: 155,776 logits per step are too big to commit -- keep the top 256 and every 16th
        top = sl.float().topk(256, dim=-1)
        out["slow_top_idx"], out["slow_top_val"] = top.indices.to(torch.int32), sl.gather(1, top.indices)
        out["slow_strided"] = sl[:, ::16].clone()
    return out


def main():
    t0 = time.time()
    torch.set_num_threads(8)
    only = set(sys.argv[1:])      # e.g. `make_golden.py projected`: (re)generate the fixtures of the named variants only
    for name, cfg in variant_configs().items():
        if only and name not in only:
            continue
        sd = make_state_dict(cfg, seed=0)
        prompt = synthetic_prompt(cfg, 5, 12, 4, seed=1)
        for mode in ("sampled", "greedy", "hot"):
            g = run_case(cfg, sd, prompt, 40, mode, seed=11)
            torch.save(g, OUT / f"tiny_{name}_{mode}.pt")
            print(name, mode, g["seq"].shape, f"{time.time() - t0:.0f}s")
        # early stop: <|im_end|> reachable
        sd2 = make_state_dict(cfg, seed=0, eos_reachable=True)
        g = run_case(cfg, sd2, prompt, 40, "sampled", seed=11)
        torch.save(g, OUT / f"tiny_{name}_eos.pt")
        print(name, "eos", g["seq"].shape, "ended with", int(g["seq"][0, -1]), "im_end", cfg.im_end_id)
    if only and "s1mini" not in only:
        return
    cfg = s1_mini_config()
    sd = make_state_dict(cfg, seed=0)
    prompt = synthetic_prompt(cfg, 6, 20, 4, seed=1)
    for mode in ("sampled", "greedy"):
        g = run_case(cfg, sd, prompt, 6, mode, seed=11, full_logits=False)
        torch.save(g, OUT / f"s1mini_{mode}.pt")
        print("s1mini", mode, g["seq"].shape, f"{time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
