#!/usr/bin/env python
"""bench.py -- dual-AR decode throughput on B200 (BASELINE.json metric: semantic tokens/s, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--torch-baselines]

Workload (BASELINE.json configs[1]): openaudio-s1-mini shape, seeded random-init weights, bs=1, a ~220-token
prompt holding one ~10 s VoiceProfile (215 frames), then K decode steps (default 1024) with the FishTTS
sampling defaults T=0.7 / top_p=0.8 / repetition_penalty=1.1.  One "step" = one decode_one_token call =
one semantic id + num_codebooks codes.

  value     device-timed tokens/s: K CUDA-graph replays of the decode step between two CUDA events, prompt
            already prefilled and every input resident in HBM (weights 1.4 GB >> 126 MB L2, so no L2 flush is needed).
  e2e       the same metric through the per-step C-ABI call a user of the reference's seam makes (dualar_step =
            decode_one_token_ar) inside the reference's own host loop (decode_n_tokens): every step copies its inputs (token
            column, position, repetition window) from pinned HOST memory, runs, and reads the sampled column back to the host;
            `e2e.request` adds one whole dualar_generate request (HOST prompt in, HOST tokens out, prefill included).
  roofline  algorithmic bytes per step (unique weights once + KV over the mean context) / mean step time,
            against the measured HBM copy peak of MEASURED_PEAKS.json.
  cpu_baseline  the oracle port of the reference step on this box's host cores (bounded sample; a reported baseline).
N > 1 (torchrun): request-level replicas, one engine per GPU, no collective on the data path ("weak" scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from fish_tts_b200.config import fish_speech_1_5_config, s1_mini_config  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402

SAMPLING = dict(temperature=0.7, top_p=0.8, repetition_penalty=1.1)
PROMPT = dict(n_text=3, n_frames=215, n_tail=5)   # 223 positions: "prefilled ~10 s VoiceProfile reference (~220-token prompt)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text())["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_port_tokens_per_s(cfg, sd, n_steps: int, prompt_len: int = 32):
    """The oracle port of the reference decode step on the host cores (the one place bench may run oracle/)."""
    from oracle import dualar_oracle as orc
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    m = orc.OracleModel.build(cfg, sd, device="cpu")
    prompt = synthetic_prompt(cfg, 3, prompt_len - 8, 5, seed=1)
    m.setup_caches(cfg.max_seq_len)
    C1 = cfg.num_codebooks + 1
    t = [torch.tensor(v, dtype=torch.float) for v in (SAMPLING["temperature"], SAMPLING["top_p"], SAMPLING["repetition_penalty"])]
    with torch.inference_mode():
        first = orc.decode_one_token_ar(m, prompt.view(1, C1, -1), torch.arange(prompt.size(1)), *t, None)
        t0 = time.perf_counter()
        orc.decode_n_tokens(m, first.view(1, C1, -1), torch.tensor([prompt.size(1)], dtype=torch.int), n_steps, *t)
        dt = time.perf_counter() - t0
    return n_steps / dt, threads, f"same weights, {prompt.size(1)}-token prompt (prefill untimed), {n_steps} decode steps, torch CPU bf16 eager, {threads} threads"


def run_reference(args, rank: int):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; /root/reference cannot travel)."""
    if rank != 0:
        return
    cfg = s1_mini_config()
    sd = make_state_dict(cfg, seed=0)
    k = max(1, min(args.steps, 8))
    t0 = time.perf_counter()
    v, threads, sample = cpu_port_tokens_per_s(cfg, sd, k)
    line = {
        "impl": "reference", "metric": "dual_ar_decode_tokens_per_s", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": k, "warmup": 0, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": workload_config(cfg, args.gpus, k),
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the decode kernel, from the committed ncu capture
    (profiles/ncu_summary.json, written from `ncu --set full`; null when no capture is committed)."""
    try:
        with open(ROOT / "profiles" / "ncu_summary.json") as f:
            d = json.load(f)
        return float(d["dram_bytes_read_per_launch"]) + float(d["dram_bytes_write_per_launch"])
    except Exception:
        return None


def workload_config(cfg, n_gpus, steps):
    name = "openaudio-s1-mini" if cfg.num_codebooks == 10 else "fish-speech-1.5 shape"
    return {"workload": f"{name} dual-AR decode, bs=1, prefilled ~10 s VoiceProfile prompt "
                        f"({sum(PROMPT.values())} positions), {steps} generated tokens, T=0.7 top_p=0.8 rp=1.1",
            "weights": "seeded random-init (fish_tts_b200.synthetic, seed 0), bf16", "prompt_len": sum(PROMPT.values()),
            "generated": steps, "l2": f"inputs ({cfg.weight_bytes()['unique_weights'] / 1e9:.2f} GB of weights per step) exceed the 126 MB L2; no flush between steps",
            "parallelism": f"{n_gpus} independent replica(s), request-level partitioning, no collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--torch-baselines", action="store_true", help="also time the oracle's torch path on this GPU (eager and torch.compile)")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--e2e-requests", type=int, default=2)
    ap.add_argument("--model", default="s1mini", choices=["s1mini", "v15"], help="openaudio-s1-mini (BASELINE configs[1], the headline) or the fish-speech 1.5 shape (configs[2])")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the decode path")
    from fish_tts_b200.engine import DualAREngine
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
    cfg = s1_mini_config() if args.model == "s1mini" else fish_speech_1_5_config()
    sd = make_state_dict(cfg, seed=0)
    eng = DualAREngine(cfg, sd, device=local, seed=1234 + rank)
    prompt = synthetic_prompt(cfg, **PROMPT, seed=1 + rank)
    T = prompt.size(1)
    K, W = args.steps, max(args.warmup, 3)
    assert T + W + K + 1 <= cfg.max_seq_len
    launches_step, launches_prefill = eng.launches_per_step()

    # ---- device-timed: K graph replays, everything resident ---------------------------------------
    eng.prefill(prompt, W + K + 1, **SAMPLING)
    eng.decode(W)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        eng.decode(K)
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    toks, fin = eng.collect()
    assert toks.shape[1] == 1 + W + K, f"EOS or limit hit early: {toks.shape[1]} columns"
    sem = toks[0]
    assert ((sem >= cfg.semantic_begin_id) & (sem <= cfg.semantic_end_id)).all(), "non-semantic id sampled"
    tmax = torch.tensor([ms], device="cuda")
    if dist:
        dist.barrier()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    value = world * K / (ms_max / 1e3)

    # ---- end to end through the per-step C-ABI call with host buffers ------------------------------------
    # The reference's decode_n_tokens loop (inference.py:171-215) on the host around dualar_step (= decode_one_token_ar):
    # every step copies ITS inputs (token column, position, 16-wide repetition window) from pinned host memory, runs the step
    # and reads the sampled column back before the host builds the next step's inputs.  Prefill untimed, like `value` and the
    # reference arm.
    rows = cfg.num_codebooks + 1
    eng.prefill(prompt, 1, **SAMPLING)
    first, _ = eng.collect()
    dev = torch.device("cuda", local)
    # one pinned staging buffer and one device buffer hold the step's three inputs back to back (token column | position |
    # window), so a step costs ONE host-to-device copy; the C-ABI call gets views into the device buffer
    n_in = rows + 1 + rows * 16
    h_in = torch.zeros(n_in, dtype=torch.int32).pin_memory()
    h_x, h_pos, h_win = h_in[:rows], h_in[rows:rows + 1], h_in[rows + 1:].view(rows, 16)
    h_x.copy_(torch.from_numpy(first[:, -1].copy()).to(torch.int32)); h_pos[0] = T
    h_out = torch.zeros((rows,), dtype=torch.int32).pin_memory()
    d_in = h_in.to(dev)
    d_x, d_pos, d_win = d_in[:rows], d_in[rows:rows + 1], d_in[rows + 1:].view(rows, 16)
    d_par = [torch.tensor(v, dtype=torch.float, device=dev) for v in (SAMPLING["temperature"], SAMPLING["top_p"], SAMPLING["repetition_penalty"])]
    h2d = h_in.numel() * 4
    d2h = h_out.numel() * 4
    # host bookkeeping on numpy views of the pinned buffers (the GPU idles while the host prepares the next step)
    import numpy as np
    np_x, np_pos, np_win, np_out = h_x.numpy(), h_pos.numpy(), h_win.numpy(), h_out.numpy()
    np_prev = np.zeros((rows, W + K + 16), dtype=np.int32)
    stream = torch.cuda.current_stream()
    t0 = 0.0
    for i in range(W + K):
        if i == W:
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            t0 = time.perf_counter()
        np_win[:] = np_prev[:, :16] if i < 16 else np_prev[:, i - 16:i]
        d_in.copy_(h_in, non_blocking=True)
        out_d = eng.step(d_x, d_pos, d_win, *d_par)
        h_out.copy_(out_d.view(-1), non_blocking=True)
        stream.synchronize()
        np_prev[:, i] = np_out
        np_x[:] = np_out; np_pos += 1
    dt_steps = time.perf_counter() - t0
    e2e_t = torch.tensor([K / dt_steps], device="cuda")   # slowest rank
    if dist:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MIN)
    # and one whole request through dualar_generate (host prompt in, host tokens out, prefill included), for the record
    req_vals = []
    for r in range(args.e2e_requests):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = eng.generate(prompt, K, **SAMPLING)
        dt = time.perf_counter() - t0
        req_vals.append(out.shape[1] / dt)
    e2e = {"value": float(e2e_t.item()) * world, "unit": "tokens/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "step": "dualar_step (decode_one_token_ar) per token: H2D token column + position + repetition window from pinned memory, step, D2H sampled column, host sync; prefill untimed",
           "request": {"value": float(min(req_vals[1:] or req_vals)), "unit": "tokens/s",
                       "what": f"dualar_generate: H2D prompt + prefill({T}) + {K} decode steps + D2H tokens (prefill runs one position per launch)"}}

    if rank == 0:
        peak, peak_src = measured_peak()
        mean_ctx = T + W + K / 2.0
        bytes_step = cfg.algorithmic_bytes_per_token(mean_ctx)
        achieved = bytes_step / (ms_max / 1e3 / K) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                "peak_source": peak_src,
                "kernel": ("mega_kernel: the whole decode step as one persistent cooperative kernel, one launch per token"
                           if launches_step == 1 else "decode-step graph (all kernels of one token)"),
                "algorithmic_bytes_per_step": bytes_step, "mean_context": mean_ctx}
        cpu_v, cores, sample = cpu_port_tokens_per_s(cfg, sd, args.cpu_steps)
        line = {
            "metric": "dual_ar_decode_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(cfg, world, K), "e2e": e2e, "gpu_launches": launches_step * K,
            "launches_per_step": launches_step, "roofline": roof,
            "cpu_baseline": {"value": cpu_v, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
            "clocks": clocks.summary(), "rtf_decode_only": (ms_max / K / 1e3) / (2048 / 44100.0),
        }
        if args.torch_baselines:
            line["torch_baselines"] = torch_baselines(cfg, sd, prompt, local)
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def torch_baselines(cfg, sd, prompt, dev):
    """The oracle's torch restatement of the reference step on this GPU: eager and torch.compile(reduce-overhead)
    (what inference.py:406-412 does).  Reported beside our number; oracle code is only ever the thing compared against."""
    from torch.nn.attention import SDPBackend, sdpa_kernel
    from oracle import dualar_oracle as orc
    out = {}
    m = orc.OracleModel.build(cfg, sd, device=f"cuda:{dev}")
    m.setup_caches(cfg.max_seq_len)
    C1 = cfg.num_codebooks + 1
    t = [torch.tensor(v, device=m.device, dtype=torch.float) for v in (0.7, 0.8, 1.1)]
    p = prompt.to(m.device)
    with torch.inference_mode():
        first = orc.decode_one_token_ar(m, p.view(1, C1, -1), torch.arange(p.size(1), device=m.device), *t, None)
        for name, n in (("eager", 24),):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            orc.decode_n_tokens(m, first.view(1, C1, -1), torch.tensor([p.size(1)], device=m.device, dtype=torch.int), n, *t)
            torch.cuda.synchronize(); out[f"torch_{name}_tokens_per_s"] = n / (time.perf_counter() - t0)
    try:
        step = torch.compile(lambda x, ip, w: orc.decode_one_token_ar(m, x, ip, *t, w), mode="reduce-overhead", fullgraph=True)
        fn = lambda x, input_pos, previous_tokens, **kw: step(x, input_pos, previous_tokens)
        with torch.inference_mode():
            for n in (8, 64):   # first call compiles
                torch.cuda.synchronize(); t0 = time.perf_counter()
                orc.decode_n_tokens(m, first.view(1, C1, -1), torch.tensor([p.size(1)], device=m.device, dtype=torch.int), n, *t,
                                    decode_one_token=fn)
                torch.cuda.synchronize(); dt = time.perf_counter() - t0
            out["torch_compile_tokens_per_s"] = 64 / dt
    except Exception as ex:   # inductor may be unusable on the box
        out["torch_compile_error"] = str(ex)[:200]
    return out


if __name__ == "__main__":
    main()
