"""ORACLE SUPPORT (test infrastructure) -- time the UNMODIFIED reference's decode path beside ours.

    python -m oracle.ref_bench --device cpu  --model s1mini --prompt-len 223 --warmup 1 --steps 8  --budget 90
    python -m oracle.ref_bench --device cuda --compile 1 ...      (what FishTTS does: torch.compile, mode="reduce-overhead")

Runs the reference's own ``init_model`` + ``generate`` (fish_tts/models/inference.py:387-414, 279-384) -- imported from
/root/reference in the build container, from oracle/_ref/ (oracle/make_ref.py) on the GPU box -- on a fabricated model directory
holding the SAME seeded random-init weights and the same prompt as bench.py's own arm, and prints one JSON line.  Only the decode
loop is timed, like our arm: the injected ``decode_one_token`` callable (the reference's seam) is wrapped with a stopwatch; the
prefill goes through ``decode_one_token_ar`` directly (inference.py:353) and is not in the timed region.  ``--budget`` bounds the
timed region in seconds: the loop is cut short by raising from the callable once the budget is spent.
bench.py runs this in a subprocess (a slow Inductor compile or a crash must not take the bench line down with it).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


class _BudgetSpent(Exception):
    pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--compile", type=int, default=0)
    ap.add_argument("--model", default="s1mini", choices=["s1mini", "v15", "tiny"])
    ap.add_argument("--prompt-len", type=int, default=223)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--budget", type=float, default=90.0)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    os.environ.setdefault("HOME", "/tmp")
    from fish_tts_b200.config import fish_speech_1_5_config, s1_mini_config, tiny_config
    from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt
    from oracle import ref_harness

    out = {"impl": "reference", "device": args.device, "compile": bool(args.compile)}
    if not ref_harness.reference_available():
        out["unavailable"] = f"reference not found under {ref_harness.REFERENCE_ROOT} (run oracle/make_ref.py where /root/reference exists)"
        print(json.dumps(out), flush=True)
        return
    threads = args.threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    cfg = {"s1mini": s1_mini_config, "v15": fish_speech_1_5_config, "tiny": tiny_config}[args.model]()
    sd = make_state_dict(cfg, seed=0)
    prompt = synthetic_prompt(cfg, 3, args.prompt_len - 8, 5, seed=1)
    work = Path(tempfile.gettempdir()) / f"dualar_ref_model_{args.model}"
    t_load = time.perf_counter()
    _, inference = ref_harness.import_reference()
    d = ref_harness.fabricate_model_dir(cfg, sd, work)
    del sd
    model, decode_one_token = inference.init_model(str(d), device=args.device, precision=torch.bfloat16, compile=bool(args.compile))
    out["load_s"] = time.perf_counter() - t_load

    stamps = []
    sync = torch.cuda.synchronize if args.device == "cuda" else (lambda: None)
    W, K = args.warmup, args.steps

    def timed_step(**kw):
        sync()
        now = time.perf_counter()
        if len(stamps) > W and now - stamps[W] > args.budget:
            raise _BudgetSpent()
        stamps.append(now)
        return decode_one_token(**kw)

    t0 = time.perf_counter()
    try:
        inference.generate(model=model, prompt=prompt.to(args.device), max_new_tokens=W + K + 1, audio_masks=None, audio_parts=None,
                           decode_one_token=timed_step, temperature=0.7, top_p=0.8, repetition_penalty=1.1)
    except _BudgetSpent:
        pass
    sync()
    t_end = time.perf_counter()
    # stamps[i] = entry of decode call i; call i lasts until stamps[i + 1] (or the end): steps W .. n-1 are timed
    n = len(stamps)
    timed = n - W
    if timed < 1:
        out["unavailable"] = f"no timed decode step fitted the budget ({n} calls, first call took {t_end - stamps[0] if stamps else float('nan'):.1f} s incl. compilation)"
    else:
        dt = t_end - stamps[W]
        out.update(value=timed / dt, unit="tokens/s", steps=timed, warmup=W, ms_per_step=1e3 * dt / timed, cores=threads,
                   first_call_s=(stamps[1] - stamps[0]) if n > 1 else None, total_s=t_end - t0,
                   sample=(f"the reference's own init_model + generate ({'torch.compile(reduce-overhead)' if args.compile else 'eager'}, {args.device}"
                           f"{', ' + str(threads) + ' threads' if args.device == 'cpu' else ''}) on the same seeded weights and the same {prompt.size(1)}-position prompt; "
                           f"{timed} decode steps timed after {W} warm-up call(s), prefill untimed"))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
