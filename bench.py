#!/usr/bin/env python
"""bench.py -- dual-AR decode throughput on B200 (BASELINE.json metric: semantic tokens/s at bs1 & bs32, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload decode|batch|utterances|stream] [--batch B] [--model s1mini|v15] [--no-extras]

--workload decode (default; BASELINE.json configs[1], the configuration the metric is quoted on): openaudio-s1-mini shape, seeded
random-init weights, bs = 1, a 223-position prompt holding one ~10 s VoiceProfile (215 frames), then K decode steps (default 1024)
with the FishTTS sampling defaults T=0.7 / top_p=0.8 / repetition_penalty=1.1.  One "step" = one decode_one_token call = one
semantic id + num_codebooks codes.

  value      device-timed tokens/s: K CUDA-graph replays of the decode step between two CUDA events, prompt already prefilled and
             every input resident in HBM (weights 1.4 GB >> 126 MB L2, so no L2 flush is needed).
  e2e        the same metric through the per-step C-ABI call a user of the reference's seam makes (dualar_step =
             decode_one_token_ar) inside the reference's own host loop (decode_n_tokens): every step copies its inputs (token column,
             position, repetition window) from pinned HOST memory, runs, and reads the sampled column back to the host;
             `e2e.request` = one whole dualar_generate request (HOST prompt in, tensor-core prefill, HOST tokens out).
  roofline   algorithmic bytes per step (unique weights once + KV over the mean context) / mean step time, against the
             measured HBM copy peak of MEASURED_PEAKS.json.
  Extras on the same line at N = 1 (skipped with --no-extras): `batch32` (configs[3]: 32 request slots, mixed prompts, tcgen05
  GEMMs), `batch128` (128 slots as 4 concurrently running groups of 32), `prefill` (223 positions through the tensor-core prefill), `streaming` (first-chunk latency / RTF through the chunked
  hand-off), `cpu_baseline` and `torch_baselines`: the UNMODIFIED reference's own init_model + generate (oracle/ref_bench.py on the
  copy oracle/make_ref.py ships) on this box's host cores and, under torch.compile(mode="reduce-overhead"), on this GPU.
--workload batch: the batched decode step alone (configs[3]); --workload utterances: configs[4], N utterances with mixed lengths
dealt longest-first to the ranks, each rank running them through its request slots with continuous batching (--batch 1: one at a
time through the batch-1 kernel); --workload stream: configs[2], one 646-token utterance through the streaming hand-off.
N > 1 (torchrun): request-level replicas, one engine per GPU, no collective on the data path ("weak" scaling).
--impl reference: the reference's CPU path on this box's host cores (rank 0 only), same metric, config and JSON line.
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
FRAME_S = 2048 / 44100.0                          # audio seconds per semantic token (vocoder.py:854, 872)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU works (B200_PROFILING.md recipe).  The sampler runs across the warm-up, the
    timed region and the end-to-end loop (seconds of continuous decode), so that even a 30 ms timed region has samples beside it;
    `sm_mhz` is the median over the samples taken under load (SM clock above the idle floor)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
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
        load = sorted(v for v in sm if v > 0.5 * mx) or sorted(sm)
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": len([v for v in sm if v > 0.5 * mx]),
                "window": "warm-up + timed region + end-to-end loop"}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text())["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the decode kernel from the committed ncu capture
    (profiles/ncu_summary.json), with the context it was captured at; (None, None) when no capture is committed."""
    try:
        d = json.loads((ROOT / "profiles" / "ncu_summary.json").read_text())
        return float(d["dram_bytes_read_per_launch"]) + float(d["dram_bytes_write_per_launch"]), d.get("context_positions")
    except Exception:
        return None, None


def model_cfg(name):
    return s1_mini_config() if name == "s1mini" else fish_speech_1_5_config()


def workload_config(cfg, n_gpus, steps, what="decode", **extra):
    name = "openaudio-s1-mini" if cfg.num_codebooks == 10 else "fish-speech-1.5 shape"
    T = sum(PROMPT.values())
    base = {"weights": "seeded random-init (fish_tts_b200.synthetic, seed 0), bf16",
            "l2": f"inputs ({cfg.weight_bytes()['unique_weights'] / 1e9:.2f} GB of weights per step) exceed the 126 MB L2; no flush between steps",
            "parallelism": f"{n_gpus} independent replica(s), request-level partitioning, no collective"}
    if what == "decode":
        base.update(workload=f"{name} dual-AR decode, bs=1, prefilled ~10 s VoiceProfile prompt ({T} positions), {steps} generated tokens, "
                             "T=0.7 top_p=0.8 rp=1.1", prompt_len=T, generated=steps)
    base.update(extra)
    return base


# ---- the reference beside us --------------------------------------------------------------------------------------------------------
def run_ref_bench(device, compile_, model, prompt_len, warmup, steps, budget, timeout):
    """oracle/ref_bench.py in a subprocess: the UNMODIFIED reference's init_model + generate.  Returns its JSON dict."""
    cmd = [sys.executable, "-m", "oracle.ref_bench", "--device", device, "--compile", str(int(compile_)), "--model", model,
           "--prompt-len", str(prompt_len), "--warmup", str(warmup), "--steps", str(steps), "--budget", str(budget)]
    try:
        r = subprocess.run(cmd, cwd=str(ROOT), capture_output=True, text=True, timeout=timeout)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"unavailable": f"no JSON from oracle.ref_bench (rc {r.returncode}): {(r.stderr or r.stdout)[-300:]}"}
    except subprocess.TimeoutExpired:
        return {"unavailable": f"oracle.ref_bench did not finish within {timeout} s"}
    except Exception as ex:
        return {"unavailable": str(ex)[:300]}


def cpu_port_tokens_per_s(cfg, n_steps: int, prompt_len: int):
    """fallback when no copy of the reference is on the box: the oracle port of the reference step on the host cores"""
    from oracle import dualar_oracle as orc
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = make_state_dict(cfg, seed=0)
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
    return n_steps / dt, threads, f"oracle port (no copy of the reference on this box), same weights, {prompt.size(1)}-position prompt (prefill untimed), {n_steps} decode steps, torch CPU bf16 eager, {threads} threads"


def cpu_baseline(args, cfg, steps):
    """the reference's CPU path on this box's host cores: bounded sample of the same workload (same weights, same 223-position prompt)"""
    T = sum(PROMPT.values())
    r = run_ref_bench("cpu", False, args.model, T, 1, steps, budget=60, timeout=420)
    if "value" in r:
        return {"value": r["value"], "unit": "tokens/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"], "steps": r["steps"]}
    v, threads, sample = cpu_port_tokens_per_s(cfg, min(steps, 6), T)
    return {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample, "reference_unavailable": r.get("unavailable")}


def run_reference(args, rank: int):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, our arm's metric / config / line."""
    if rank != 0:
        return
    cfg = model_cfg(args.model)
    t0 = time.perf_counter()
    k = max(1, min(args.steps, 64))
    w = max(1, min(args.warmup, 2))
    T = sum(PROMPT.values())
    r = run_ref_bench("cpu", False, args.model, T, w, k, budget=120, timeout=600)
    if "value" in r:
        v, steps_done = r["value"], r["steps"]
        cb = {"value": v, "unit": "tokens/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"]}
    else:
        v, threads, sample = cpu_port_tokens_per_s(cfg, min(k, 6), T)
        steps_done = min(k, 6)
        cb = {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample, "reference_unavailable": r.get("unavailable")}
    line = {
        "impl": "reference", "metric": "dual_ar_decode_tokens_per_s", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": steps_done, "warmup": w, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(cfg, args.gpus, steps_done, same_config=True,
                                  note=f"bounded sample: {steps_done} of the requested {args.steps} decode steps (the CPU path runs at ~2 tokens/s)"),
        "cpu_baseline": cb, "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ---- measurements on our engine -----------------------------------------------------------------------------------------------------
def measure_batch(eng, cfg, B, steps, warmup, slot_len=1152, seed=2, group_slots=None):
    """configs[3]: B request slots, prompt lengths uniform in [64, 512], one batched step = one token for every slot"""
    import numpy as np
    eng.batch_init(B, slot_len, group_slots)
    groups = int(eng.batch_read("groups")[0])
    rng = np.random.default_rng(seed)
    lens = rng.integers(64, 513, size=B)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for sl in range(B):
        eng.batch_prefill(sl, synthetic_prompt(cfg, 3, int(lens[sl]) - 8, 5, seed=10 + sl), warmup + steps + 8, **SAMPLING, seed=100 + sl)
    e1.record(); torch.cuda.synchronize()
    t_pf = e0.elapsed_time(e1)
    eng.batch_decode(warmup); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.batch_decode(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n_gen = eng.batch_read("n_gen")
    assert int(n_gen.min()) == warmup + steps, f"a slot stopped early: {n_gen.tolist()}"
    toks = eng.batch_read("tokens")[:, 0]
    assert ((toks >= cfg.semantic_begin_id) & (toks <= cfg.semantic_end_id)).all(), "non-semantic id sampled"
    ctx = float(lens.mean()) + warmup + steps / 2.0
    wb = cfg.weight_bytes()
    by = wb["unique_weights"] + wb["kv_per_pos"] * (ctx + 1) * B
    peak, peak_src = measured_peak()
    launches = int(eng.batch_read("launches")[0])
    return {"value": B / ms * 1e3, "unit": "tokens/s", "batch": B, "groups": groups, "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "launches_per_step": launches, "prefill_ms_total": t_pf, "mean_prompt": float(lens.mean()),
            "workload": f"batched decode bs={B}" + (f" as {groups} concurrent groups of {-(-B // groups)} slots" if groups > 1 else "") +
                        f", prompt lengths uniform in [64, 512] (seed {seed}), one KV cache per slot, T=0.7 top_p=0.8 rp=1.1",
            "roofline": {"bound": "hbm", "achieved": by / ms / 1e6, "peak": peak, "unit": "GB/s", "frac": by / ms / 1e6 / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": by, "mean_context": ctx,
                         "kernel": "one batched step: tcgen05 GEMMs (weights streamed once per group of slots) + per-slot attention / samplers; "
                                   "algorithmic bytes count the weights ONCE per step however many groups stream them"}}


def measure_stream(eng, cfg, prompt, n_tokens, first_chunk=10, chunk=20):
    """configs[2] mechanics: first-chunk latency (prefill + first_chunk columns on the host) and RTF of the dual-AR stage through
    DualAREngine.stream (chunked copies into pinned buffers behind events, next chunk enqueued before the host waits)"""
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_first, n = None, 0
    for cols in eng.stream(prompt, n_tokens, **SAMPLING, first_chunk=first_chunk, chunk=chunk):
        if t_first is None:
            t_first = time.perf_counter() - t0
        n += cols.shape[1]
    dt = time.perf_counter() - t0
    return {"first_chunk_ms": 1e3 * t_first, "tokens": n, "tokens_per_s": n / dt, "rtf_decode_only": dt / (n * FRAME_S),
            "chunking": f"first {first_chunk} then {chunk} columns per chunk (synthesize_stream defaults, synthesizer.py:487-488)",
            "codec": "not included: the codec / vocoder stays on the reference torch path and `dac` / `audiotools` are not installed; "
                     "see `codec_stand_in`"}


def codec_stand_in(dev, n_frames=20):
    """The codec stage timed SEPARATELY with a stated stand-in (north_star: 'codec decode ... timed separately').  The reference's DAC
    (vocoder.py:824-928) cannot be built here (`dac`, `audiotools` un-vendored), so the stand-in is a torch
    transposed-convolution stack with the DAC decoder's rates [8, 8, 4, 2] and width 1536 -> 96 (synthesizer.py:255-258) in bf16:
    the same frames-in / 2048-samples-per-frame-out shape of work, NOT the same network."""
    import torch.nn as nn
    layers, ch = [nn.Conv1d(1024, 1536, 7, padding=3)], 1536
    for r in (8, 8, 4, 2):
        layers += [nn.ConvTranspose1d(ch, ch // 2, 2 * r, stride=r, padding=r // 2), nn.Conv1d(ch // 2, ch // 2, 7, padding=3)]
        ch //= 2
    layers.append(nn.Conv1d(ch, 1, 7, padding=3))
    net = nn.Sequential(*layers).to(device=dev, dtype=torch.bfloat16).eval()
    x = torch.randn(1, 1024, 4 * n_frames, device=dev, dtype=torch.bfloat16)      # 4 latent frames per token (hop 512, frame 2048)
    with torch.inference_mode():
        for _ in range(3):
            y = net(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            y = net(x)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
    return {"ms_per_chunk": 1e3 * dt, "frames": n_frames, "samples_out": int(y.shape[-1]), "rtf": dt / (n_frames * FRAME_S),
            "what": "stand-in transposed-conv decoder with the DAC decoder's rates and widths (torch, bf16), not the reference network"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="decode", choices=["decode", "batch", "utterances", "stream"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--group-slots", type=int, default=None, help="request slots per concurrently running group (default 32)")
    ap.add_argument("--utterances", type=int, default=4096)
    ap.add_argument("--no-extras", action="store_true", help="decode workload: skip batch32 / batch128 / prefill / streaming / reference baselines")
    ap.add_argument("--cpu-steps", type=int, default=24)
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
    cfg = model_cfg(args.model)
    sd = make_state_dict(cfg, seed=0)
    eng = DualAREngine(cfg, sd, device=local, seed=1234 + rank)
    if args.workload == "batch":
        return run_batch_workload(args, cfg, eng, dist, rank, world, local)
    if args.workload == "utterances":
        return run_utterances(args, cfg, eng, dist, rank, world, local)
    if args.workload == "stream":
        return run_stream_workload(args, cfg, eng, rank, local)

    prompt = synthetic_prompt(cfg, **PROMPT, seed=1 + rank)
    T = prompt.size(1)
    K, W = args.steps, max(args.warmup, 3)
    assert T + W + K + 1 <= cfg.max_seq_len
    launches_step, launches_prefill = eng.launches_per_step()

    with ClockSampler(local) as clocks:
        # ---- device-timed: K graph replays, everything resident ---------------------------------------
        eng.prefill(prompt, W + K + 1, **SAMPLING)
        eng.decode(W)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
        import numpy as np
        rows = cfg.num_codebooks + 1
        eng.prefill(prompt, 1, **SAMPLING)
        first, _ = eng.collect()
        dev = torch.device("cuda", local)
        n_in = rows + 1 + rows * 16
        h_in = torch.zeros(n_in, dtype=torch.int32).pin_memory()
        h_x, h_pos, h_win = h_in[:rows], h_in[rows:rows + 1], h_in[rows + 1:].view(rows, 16)
        h_x.copy_(torch.from_numpy(first[:, -1].copy()).to(torch.int32)); h_pos[0] = T
        h_out = torch.zeros((rows,), dtype=torch.int32).pin_memory()
        d_in = h_in.to(dev)
        d_x, d_pos, d_win = d_in[:rows], d_in[rows:rows + 1], d_in[rows + 1:].view(rows, 16)
        d_par = [torch.tensor(v, dtype=torch.float, device=dev) for v in (SAMPLING["temperature"], SAMPLING["top_p"], SAMPLING["repetition_penalty"])]
        h2d, d2h = h_in.numel() * 4, h_out.numel() * 4
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
    # one whole request through dualar_generate (host prompt in, host tokens out, tensor-core prefill included), for the record
    req_vals = []
    for r in range(args.e2e_requests):
        eng.set_option("prefix_reuse", 0)      # every request pays its whole prefill here
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = eng.generate(prompt, K, **SAMPLING)
        dt = time.perf_counter() - t0
        req_vals.append(out.shape[1] / dt)
    eng.set_option("prefix_reuse", 1)
    e2e = {"value": float(e2e_t.item()) * world, "unit": "tokens/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "step": "dualar_step (decode_one_token_ar) per token: H2D token column + position + repetition window from pinned memory, step, D2H sampled column, host sync; prefill untimed",
           "request": {"value": float(min(req_vals[1:] or req_vals)), "unit": "tokens/s",
                       "what": f"dualar_generate: H2D prompt + tensor-core prefill({T} positions) + {K} decode steps + D2H tokens"}}

    if rank == 0:
        peak, peak_src = measured_peak()
        mean_ctx = T + W + K / 2.0
        bytes_step = cfg.algorithmic_bytes_per_token(mean_ctx)
        achieved = bytes_step / (ms_max / 1e3 / K) / 1e9
        traffic, traffic_ctx = ncu_traffic()
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_context": (f"ncu capture at context ~{traffic_ctx} positions (profiles/ncu_summary.json); this line ran at mean context {mean_ctx:.0f}"
                                    if traffic is not None else None),
                "peak_source": peak_src,
                "kernel": ("mega_kernel: the whole decode step as one persistent cooperative kernel, one launch per token"
                           if launches_step == 1 else "decode-step graph (all kernels of one token)"),
                "algorithmic_bytes_per_step": bytes_step, "mean_context": mean_ctx}
        line = {
            "metric": "dual_ar_decode_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(cfg, world, K), "e2e": e2e, "gpu_launches": launches_step * K,
            "launches_per_step": launches_step, "roofline": roof, "clocks": clocks.summary(), "rtf_decode_only": (ms_max / K / 1e3) / FRAME_S,
        }
        if world == 1 and not args.no_extras:
            # ---- the other rows of the metric, same process, same weights ------------------------------------------------------
            prefill = {}
            for mode, name in ((0, "tensor_core_ms"), (1, "one_position_per_launch_ms")):
                eng.set_option("prefill_mode", mode); eng.set_option("prefix_reuse", 0)
                eng.prefill(prompt, 2, **SAMPLING); torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(); eng.prefill(prompt, 2, **SAMPLING); a1.record(); torch.cuda.synchronize()
                prefill[name] = a0.elapsed_time(a1)
            eng.set_option("prefill_mode", 0); eng.set_option("prefix_reuse", 1)
            prefill.update(positions=T, kernels=int(eng.read("prefill_launches")[0]),
                           what="dualar_prefill of the 223-position prompt incl. the decode step that samples the first token; tensor_core = positions [0, T-1) through the tcgen05 GEMMs")
            line["prefill"] = prefill
            line["streaming"] = measure_stream(eng, cfg, prompt, 646)
            line["streaming"]["codec_stand_in"] = codec_stand_in(torch.device("cuda", local))
            line["batch32"] = measure_batch(eng, cfg, 32, 128, 16)
            eng.close()
            # 128 request slots as 4 concurrently running groups of 32 (a fresh engine: slots are allocated once per engine)
            eng = DualAREngine(cfg, sd, device=local, seed=1234 + rank)
            line["batch128"] = measure_batch(eng, cfg, 128, 64, 8)
            eng.close()
            torch.cuda.empty_cache()
            line["cpu_baseline"] = cpu_baseline(args, cfg, args.cpu_steps)
            tb = run_ref_bench("cuda", True, args.model, T, 4, 128, budget=60, timeout=420)
            line["torch_baselines"] = ({"torch_compile_tokens_per_s": tb["value"], "first_call_s": tb.get("first_call_s"), "load_s": tb.get("load_s"),
                                        "what": tb["sample"]} if "value" in tb else {"unavailable": tb.get("unavailable")})
        else:
            line["cpu_baseline"] = None if world > 1 else cpu_baseline(args, cfg, args.cpu_steps)
            if world > 1:
                line["cpu_baseline_note"] = "measured at N = 1 only (rank 0's host cores are shared with the other ranks' launch threads at N > 1)"
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def run_batch_workload(args, cfg, eng, dist, rank, world, local):
    with ClockSampler(local) as clocks:
        m = measure_batch(eng, cfg, args.batch, args.steps if args.steps != 1024 else 256, max(args.warmup, 3), seed=2 + rank, group_slots=args.group_slots)
    t = torch.tensor([m["ms_per_step"]], device="cuda")
    if dist:
        dist.barrier(); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        line = {"metric": "dual_ar_decode_tokens_per_s", "value": world * args.batch / ms * 1e3, "unit": "tokens/s", "n_gpus": world, "steps": m["steps"],
                "warmup": m["warmup"], "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(cfg, world, m["steps"], what="batch", workload=m["workload"], batch=args.batch),
                "gpu_launches": m["launches_per_step"] * m["steps"], "launches_per_step": m["launches_per_step"], "roofline": m["roofline"],
                "clocks": clocks.summary(), "prefill_ms_total": m["prefill_ms_total"]}
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier(); dist.destroy_process_group()


def run_utterances(args, cfg, eng, dist, rank, world, local):
    """configs[4]: N synthetic utterances (prompt lengths uniform in [64, 512], target lengths uniform in [128, 1024], EOS forced at
    the target length), dealt longest-first to the ranks (fish_tts_b200.replicas); value = total tokens / max rank seconds, prefill included"""
    from fish_tts_b200 import replicas
    utts = replicas.synthetic_utterances(cfg, args.utterances)
    B = args.batch
    sync = torch.cuda.synchronize
    if dist:
        dist.barrier()
    with ClockSampler(local) as clocks:
        if B > 1:
            eng.batch_init(B, 512 + 1024 + 64, args.group_slots)
            res = replicas.run_rank_batched(eng, utts, rank, world, B, sync=sync)
        else:
            res = replicas.run_rank(lambda u: eng.generate(u.prompt, u.max_new_tokens, u.temperature, u.top_p, u.repetition_penalty), utts, rank, world, sync=sync)
    agg = replicas.aggregate(res, dist, torch.device("cuda", local))
    secs = torch.tensor([res.seconds], device="cuda", dtype=torch.float64)
    gathered = [torch.zeros_like(secs) for _ in range(world)] if dist else [secs]
    if dist:
        dist.all_gather(gathered, secs)
    if rank == 0:
        per_rank = [float(g.item()) for g in gathered]
        line = {"metric": "dual_ar_decode_tokens_per_s", "value": agg["tokens_per_s"], "unit": "tokens/s", "n_gpus": world, "steps": agg["tokens"], "warmup": 0,
                "ms_per_step": 1e3 * agg["seconds"] / max(agg["tokens"], 1) * world, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": workload_config(cfg, world, agg["tokens"], what="utterances",
                                          workload=f"{args.utterances} synthetic utterances, prompts uniform in [64, 512], targets uniform in [128, 1024] tokens (seed 3), "
                                                   f"{'continuous batching over ' + str(B) + ' request slots per GPU' if B > 1 else 'one request at a time per GPU (batch-1 kernel)'}, "
                                                   "prefill included, longest-first partition"),
                "total_tokens": agg["tokens"], "seconds_max_rank": agg["seconds"], "seconds_per_rank": per_rank,
                "host_blocked_on_gpu_s_rank0": getattr(res, "wait_seconds", None),
                "imbalance": (max(per_rank) - min(per_rank)) / max(per_rank) if per_rank else 0.0, "clocks": clocks.summary()}
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier(); dist.destroy_process_group()


def run_stream_workload(args, cfg, eng, rank, local):
    """configs[2]: one 30 s utterance (646 tokens) through the streaming hand-off, chunk_tokens=20, min_first_chunk=10"""
    prompt = synthetic_prompt(cfg, **PROMPT, seed=1)
    measure_stream(eng, cfg, prompt, 64)      # warm
    with ClockSampler(local) as clocks:
        m = measure_stream(eng, cfg, prompt, 646)
    m["codec_stand_in"] = codec_stand_in(torch.device("cuda", local))
    if rank == 0:
        line = {"metric": "dual_ar_decode_tokens_per_s", "value": m["tokens_per_s"], "unit": "tokens/s", "n_gpus": 1, "steps": m["tokens"], "warmup": 64,
                "ms_per_step": 1e3 / m["tokens_per_s"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(cfg, 1, m["tokens"], what="stream",
                                          workload=f"{'fish-speech-1.5 shape' if cfg.num_codebooks != 10 else 'openaudio-s1-mini'} streaming synthesis of a 30 s utterance "
                                                   "(646 tokens), 223-position prompt, chunk_tokens=20, min_first_chunk=10, prefill included"),
                "streaming": m, "clocks": clocks.summary()}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
