"""Request-level partitioning of independent utterances across the GPUs of one box.

The decode step does not shard: a 0.7 B-parameter batch-1 step is latency/bandwidth bound and tensor parallelism
would add two all-reduces per layer to a sub-millisecond step.  Multi-GPU is therefore REPLICAS ONLY: one process per
GPU (the reference's ``get_instance`` is a per-process singleton without a device index, synthesizer.py:661-710), each
owning a full engine; utterances are dealt out longest-first; nothing crosses NVLink on the data path.
``torch.distributed`` is used for the bookkeeping around it only (barrier, max-over-ranks time, token totals) -- NCCL
on the GPU box, gloo in the CPU tests.
"""

from __future__ import annotations

import heapq
import time
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np


@dataclass
class Utterance:
    uid: int
    prompt: np.ndarray            # (num_codebooks + 1, T) int32
    max_new_tokens: int
    temperature: float = 0.7
    top_p: float = 0.8
    repetition_penalty: float = 1.1

    @property
    def cost(self) -> float:
        """decode steps dominate; prefill positions cost about a third of a decode step each"""
        return self.max_new_tokens + self.prompt.shape[1] / 3.0


def partition_longest_first(costs: Sequence[float], world: int) -> list[list[int]]:
    """LPT greedy: sort by cost descending, always give the next item to the least-loaded rank.
    Deterministic (ties by index), every item assigned exactly once; makespan <= 4/3 OPT."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0.0, r) for r in range(world)]
    heapq.heapify(heap)
    out = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(i)
        heapq.heappush(heap, (load + costs[i], r))
    return out


@dataclass
class RankResult:
    rank: int
    uids: list = field(default_factory=list)
    codes: dict = field(default_factory=dict)     # uid -> (num_codebooks + 1, n) int32
    tokens: int = 0
    seconds: float = 0.0
    wait_seconds: float = 0.0


def run_rank(generate_fn: Callable[[Utterance], np.ndarray], utterances: Sequence[Utterance], rank: int, world: int,
             sync: Optional[Callable[[], None]] = None) -> RankResult:
    """Run this rank's share. ``generate_fn`` is ``lambda u: engine.generate(u.prompt, u.max_new_tokens, ...)``;
    ``sync`` (e.g. torch.cuda.synchronize) brackets the timed region."""
    mine = partition_longest_first([u.cost for u in utterances], world)[rank]
    res = RankResult(rank=rank)
    if sync:
        sync()
    t0 = time.perf_counter()
    for i in mine:
        u = utterances[i]
        out = generate_fn(u)
        res.uids.append(u.uid)
        res.codes[u.uid] = out
        res.tokens += int(out.shape[1])
    if sync:
        sync()
    res.seconds = time.perf_counter() - t0
    return res


def run_rank_batched(engine, utterances: Sequence[Utterance], rank: int, world: int, max_batch: int, poll_steps: int = 16,
                     sync: Optional[Callable[[], None]] = None, seed_base: int = 0, pipeline: Optional[bool] = None) -> RankResult:
    """This rank's share through the engine's request slots with CONTINUOUS BATCHING: every free slot is refilled from the queue
    (tensor-core prefill), ``poll_steps`` batched decode steps run without a host round trip, finished slots are collected and
    released.  ``engine`` is a DualAREngine on which ``batch_init(max_batch, ...)`` has been called.  Longest-first within the
    rank keeps the tail short.  One device -> host read of the ``done`` flags per ``poll_steps`` steps.

    ``pipeline`` (default: on when the engine runs several request groups): as soon as the ``done`` flags of burst k are on the host,
    burst k+1 is enqueued, and only THEN are the finished requests collected, their slots released and refilled -- the host work
    (D2H of the codes, ~230 kernel launches per prefill) overlaps the GPU's next burst instead of leaving it idle.  A refilled slot
    joins at burst k+2; its prefill runs beside burst k+1 on the engine's prefill stream (finished requests park their slot, so the
    steps in flight do not touch the cache being filled)."""
    mine = partition_longest_first([u.cost for u in utterances], world)[rank]
    queue = sorted(mine, key=lambda i: (-utterances[i].cost, i))
    res = RankResult(rank=rank)
    slots: dict[int, int] = {}          # slot -> utterance index
    free = list(range(max_batch))
    if pipeline is None:
        pipeline = int(engine.batch_read("groups")[0]) > 1
    if pipeline and hasattr(engine, "set_option"):
        engine.set_option("batch_decode_join", 0)      # batch_read / batch_collect wait for what they need; decode itself returns at once

    def refill():
        while free and queue:
            i = queue.pop(0)
            u = utterances[i]
            sl = free.pop()
            engine.batch_prefill(sl, u.prompt, u.max_new_tokens, u.temperature, u.top_p, u.repetition_penalty, seed=seed_base + u.uid)
            slots[sl] = i

    if sync:
        sync()
    t0 = time.perf_counter()
    t_wait = 0.0
    in_flight = False
    refill()
    while queue or slots:
        if not in_flight:
            engine.batch_decode(poll_steps)
        tw = time.perf_counter()
        done = engine.batch_read("done").numpy()          # waits for every burst enqueued so far
        t_wait += time.perf_counter() - tw
        finished = [s for s in slots if done[s]]
        in_flight = pipeline and len(slots) > len(finished)
        if in_flight:
            engine.batch_decode(poll_steps)               # the GPU goes on while the host collects and refills
        for sl in finished:
            u = utterances[slots.pop(sl)]
            out, fin = engine.batch_collect(sl)
            assert fin
            engine.batch_release(sl)
            free.append(sl)
            res.uids.append(u.uid)
            res.codes[u.uid] = out
            res.tokens += int(out.shape[1])
        refill()
    if pipeline and hasattr(engine, "set_option"):
        engine.set_option("batch_decode_join", 1)
    if sync:
        sync()
    res.seconds = time.perf_counter() - t0
    res.wait_seconds = t_wait          # host blocked on the GPU (the rest of `seconds` is host work: collects, releases, prefill enqueues)
    return res


def aggregate(res: RankResult, dist=None, device=None) -> dict:
    """Whole-job numbers: total tokens over the MAX rank time (never a sum of per-rank rates)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return {"tokens": res.tokens, "seconds": res.seconds, "tokens_per_s": res.tokens / max(res.seconds, 1e-9), "world": 1}
    t = torch.tensor([float(res.tokens)], device=device, dtype=torch.float64)
    s = torch.tensor([res.seconds], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
    return {"tokens": int(t.item()), "seconds": float(s.item()), "tokens_per_s": float(t.item() / max(s.item(), 1e-9)),
            "world": dist.get_world_size()}


def synthetic_utterances(cfg, n: int, seed: int = 3, prompt_range=(64, 512), target_range=(128, 1024)) -> list[Utterance]:
    """BASELINE.json configs[4]: mixed prompt lengths and target lengths, EOS forced at the target length
    (the conditioned checkpoint never emits <|im_end|>, so ``max_new_tokens`` is the target length)."""
    from .synthetic import synthetic_prompt
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        T = int(rng.integers(prompt_range[0], prompt_range[1] + 1))
        tgt = int(rng.integers(target_range[0], target_range[1] + 1))
        prompt = synthetic_prompt(cfg, 3, T - 8, 5, seed=1000 + i).numpy()
        out.append(Utterance(uid=i, prompt=prompt, max_new_tokens=tgt))
    return out
