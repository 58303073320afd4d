"""DualAREngine -- the Python handle on one libdualar engine (one request at a time, one GPU).

PyTorch is plumbing here: it owns host/device buffers and the current CUDA stream; all compute of
the decode path happens inside libdualar.so.  There is no fallback path.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import capi
from .config import DualARConfig


def precompute_freqs_cis(seq_len: int, n_elem: int, base: float = 10000) -> torch.Tensor:
    """The RoPE table exactly as the reference builds it (fish_tts/models/llama.py:594-603):
    computed with the same torch ops so the bf16 table is bit-identical to the model buffer."""
    freqs = 1.0 / (base ** (torch.arange(0, n_elem, 2)[: (n_elem // 2)].float() / n_elem))
    t = torch.arange(seq_len, device=freqs.device)
    freqs = torch.outer(t, freqs)
    freqs_cis = torch.polar(torch.ones_like(freqs), freqs)
    cache = torch.stack([freqs_cis.real, freqs_cis.imag], dim=-1)
    return cache.to(dtype=torch.bfloat16)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class DualAREngine:
    """One DualARTransformer instance resident on one B200.

    ``state_dict`` uses the reference checkpoint's keys (llama.py:349-359, 511-535); tensors may live
    on the host or on the engine's device and are converted to bf16.  ``kv`` optionally binds
    caller-owned KV caches ``{"slow": [(k, v), ...], "fast": [(k, v), ...]}`` laid out like the
    reference's ``KVCache`` buffers -- that is how the step runs beneath an unmodified reference
    prefill (phase A).
    """

    def __init__(self, cfg: DualARConfig, state_dict: dict, device: int | str | torch.device = 0,
                 kv: Optional[dict] = None, seed: int = 0, freqs_cis: Optional[torch.Tensor] = None,
                 fast_freqs_cis: Optional[torch.Tensor] = None, options: Optional[dict] = None):
        self.lib = capi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("DualAREngine needs a CUDA device (sm_100a); there is no CPU path")
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise RuntimeError(f"DualAREngine cannot run on {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.cfg = cfg
        self.rows = cfg.num_codebooks + 1
        self.fast_vocab = min(1024, cfg.codebook_size)
        self.noise_per_step = cfg.vocab_size + (cfg.num_codebooks - 1) * self.fast_vocab
        self._h = C.c_void_p()
        ccfg = capi.make_config(cfg)
        capi.check(self.lib.dualar_create(C.byref(ccfg), self.device.index, C.byref(self._h)))
        try:
            tables = {
                "freqs_cis": freqs_cis if freqs_cis is not None else
                precompute_freqs_cis(cfg.max_seq_len, cfg.head_dim, cfg.rope_base),
                "fast_freqs_cis": fast_freqs_cis if fast_freqs_cis is not None else
                precompute_freqs_cis(cfg.num_codebooks, cfg.fast_head_dim, cfg.rope_base),
            }
            for key, t in list(state_dict.items()) + list(tables.items()):
                if key.endswith("k_cache") or key.endswith("v_cache") or key in ("causal_mask",):
                    continue
                if key in ("freqs_cis", "fast_freqs_cis") and key in state_dict and t is not state_dict[key]:
                    continue
                t = t.detach()
                if t.dtype != torch.bfloat16:
                    t = t.to(torch.bfloat16)
                t = t.contiguous()
                on_dev = 1 if t.is_cuda else 0
                if t.is_cuda and t.device != self.device:
                    t, on_dev = t.cpu(), 0
                capi.check(self.lib.dualar_load_weight(self._h, key.encode(), t.data_ptr(), t.numel(), on_dev))
            self._kv_keepalive = kv
            if kv is not None:
                for is_fast, name in ((0, "slow"), (1, "fast")):
                    for i, (k, v) in enumerate(kv[name]):
                        assert k.is_cuda and k.dtype == torch.bfloat16 and k.is_contiguous() and v.is_contiguous()
                        capi.check(self.lib.dualar_bind_kv(self._h, is_fast, i, k.data_ptr(), v.data_ptr()))
            capi.check(self.lib.dualar_seed(self._h, seed))
            for name, value in (options or {}).items():
                capi.check(self.lib.dualar_set_option(self._h, name.encode(), float(value)))
            capi.check(self.lib.dualar_finalize(self._h))
        except Exception:
            self.close()
            raise
        self._out = torch.zeros(self.rows, dtype=torch.int32, device=self.device)
        self._noise_keepalive = None

    # ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.dualar_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- the step: decode_one_token_ar (inference.py:83-155) -----------------------------------------
    def step(self, x: torch.Tensor, input_pos: torch.Tensor, previous_tokens: Optional[torch.Tensor],
             temperature: torch.Tensor, top_p: torch.Tensor, repetition_penalty: torch.Tensor,
             noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """All arguments are CUDA tensors as the reference passes them (x (1,C+1,1) int32, input_pos (1,)
        int32, previous_tokens a (C+1,16) int32 view, three 0-dim fp32 tensors).  Returns a (C+1, 1) int32
        view of an engine-owned buffer -- the caller clones it, as the reference's loop does."""
        xs = x.reshape(-1)
        if xs.dtype != torch.int32:
            xs = xs.to(torch.int32)
        if not xs.is_contiguous():
            xs = xs.contiguous()
        if input_pos.dtype != torch.int32:
            input_pos = input_pos.to(torch.int32)
        prev_ptr, prev_stride = None, 0
        if previous_tokens is not None:
            pt = previous_tokens
            if pt.dtype != torch.int32 or pt.stride(1) != 1 or pt.shape != (self.rows, 16):
                pt = pt.to(torch.int32).contiguous()
            prev_ptr, prev_stride = pt.data_ptr(), pt.stride(0)
        if noise is not None:
            assert noise.dtype == torch.bfloat16 and noise.is_cuda and noise.numel() >= self.noise_per_step
        capi.check(self.lib.dualar_step(self._h, xs.data_ptr(), input_pos.data_ptr(), prev_ptr, prev_stride,
                                        temperature.data_ptr(), top_p.data_ptr(), repetition_penalty.data_ptr(),
                                        _ptr(noise), self._out.data_ptr(), self._stream()))
        return self._out.view(self.rows, 1)

    # ---- the loop: generate (inference.py:279-384) ------------------------------------------------------
    def prefill(self, prompt, max_new_tokens: int, temperature: float = 0.7, top_p: float = 0.7,
                repetition_penalty: float = 1.5):
        p = np.ascontiguousarray(np.asarray(prompt.cpu() if isinstance(prompt, torch.Tensor) else prompt, dtype=np.int32))
        assert p.ndim == 2 and p.shape[0] == self.rows, f"prompt must be ({self.rows}, T)"
        self._T = p.shape[1]
        capi.check(self.lib.dualar_prefill(self._h, p.ctypes.data, p.shape[1], int(max_new_tokens), float(temperature),
                                           float(top_p), float(repetition_penalty), self._stream()))

    def decode(self, n_steps: int):
        capi.check(self.lib.dualar_decode(self._h, int(n_steps), self._stream()))

    def stream(self, prompt, max_new_tokens: int, temperature: float = 0.7, top_p: float = 0.7, repetition_penalty: float = 1.5,
               first_chunk: int = 10, chunk: int = 20):
        """Generator over host arrays (C+1, n) of freshly generated columns -- the hand-off ``synthesize_stream`` needs
        (synthesizer.py:548-559: first ``min_first_chunk`` columns, then ``chunk_tokens`` at a time).

        Each chunk is ``dualar_decode_async``: decode steps, then an asynchronous copy of the new columns into a pinned host buffer
        and an event, all on the decode stream; chunk k + 1 is enqueued BEFORE the host waits for chunk k, so the GPU keeps decoding
        while the caller (the vocoder thread) consumes a chunk.  No blocking ``collect()`` on the decode stream."""
        self.prefill(prompt, max_new_tokens, temperature, top_p, repetition_penalty)
        limit = int(max_new_tokens) if max_new_tokens and max_new_tokens > 0 else self.cfg.max_seq_len
        size = max(first_chunk, chunk) + 1
        bufs = [torch.zeros((self.rows, size), dtype=torch.int32).pin_memory() for _ in range(3)]
        states = [torch.zeros(5, dtype=torch.int32).pin_memory() for _ in range(3)]
        pending = []      # (ticket, buffer index, pitch)

        def enqueue(n_steps, k):
            t = C.c_int(0)
            capi.check(self.lib.dualar_decode_async(self._h, int(n_steps), bufs[k].data_ptr(), states[k].data_ptr(), self._stream(), C.byref(t)))
            pending.append((t.value, k, n_steps + 1))

        enqueue(first_chunk - 1, 0)          # the prefill call already produced column 0
        k, sent, enq = 1, 0, first_chunk
        if enq < limit:
            enqueue(min(chunk, limit - enq), k); enq += min(chunk, limit - enq); k = (k + 1) % 3
        while pending:
            ticket, bi, pitch = pending.pop(0)
            capi.check(self.lib.dualar_wait(self._h, ticket, 1))
            n_gen, done, err, first, ncopied = (int(v) for v in states[bi])
            if err:
                raise capi.DualarError(-5, f"device fault flag {err} during streaming decode")
            n_valid = max(0, min(n_gen, first + ncopied) - first)
            if n_valid:
                flat = bufs[bi].view(-1)[: self.rows * pitch].view(self.rows, pitch)
                yield flat[:, :n_valid].numpy().copy()
                sent += n_valid
            if done:
                for tk, _, _ in pending:      # chunks enqueued past the end are no-ops on the device; drain their events
                    capi.check(self.lib.dualar_wait(self._h, tk, 1))
                return
            if enq < limit and len(pending) < 2:
                enqueue(min(chunk, limit - enq), k); enq += min(chunk, limit - enq); k = (k + 1) % 3

    def collect(self, capacity: Optional[int] = None):
        """-> (tokens (C+1, n) int32 ndarray, finished)."""
        cap = capacity or self.cfg.max_seq_len
        out = np.zeros((self.rows, cap), dtype=np.int32)
        n, fin = C.c_int(0), C.c_int(0)
        capi.check(self.lib.dualar_collect(self._h, out.ctypes.data, cap, C.byref(n), C.byref(fin), self._stream()))
        return out[:, : n.value].copy(), bool(fin.value)

    def generate(self, prompt, max_new_tokens: int, temperature: float = 0.7, top_p: float = 0.7,
                 repetition_penalty: float = 1.5) -> np.ndarray:
        """prompt (C+1, T) int32 on the HOST -> generated columns (C+1, n) int32 on the HOST
        (the reference's ``generate`` returns prompt + these; its caller drops the last column)."""
        p = np.ascontiguousarray(np.asarray(prompt.cpu() if isinstance(prompt, torch.Tensor) else prompt, dtype=np.int32))
        assert p.ndim == 2 and p.shape[0] == self.rows, f"prompt must be ({self.rows}, T)"
        cap = self.cfg.max_seq_len
        out = np.zeros((self.rows, cap), dtype=np.int32)
        n = C.c_int(0)
        capi.check(self.lib.dualar_generate(self._h, p.ctypes.data, p.shape[1], int(max_new_tokens), float(temperature),
                                            float(top_p), float(repetition_penalty), out.ctypes.data, cap, C.byref(n),
                                            self._stream()))
        return out[:, : n.value].copy()

    # ---- batched decode: B requests share every weight byte (include/dualar.h "batched decode") -------------------
    def batch_init(self, max_batch: int, slot_seq_len: int = 0, group_slots: Optional[int] = None):
        """Allocate ``max_batch`` request slots (own KV cache of ``slot_seq_len`` positions each) and capture the batched step.
        Slots are organised in groups of ``group_slots`` (default 32): one batched step per group, the groups of a step run
        concurrently on their own streams (the step of one group is a latency-bound chain that leaves the GPU mostly idle)."""
        if group_slots is not None:
            self.set_option("batch_group_slots", group_slots)
        capi.check(self.lib.dualar_batch_init(self._h, int(max_batch), int(slot_seq_len)))
        self.max_batch = int(max_batch)
        self._batch_noise = {}

    def batch_prefill(self, slot: int, prompt, max_new_tokens: int, temperature: float = 0.7, top_p: float = 0.7,
                      repetition_penalty: float = 1.5, seed: int = 0, noise: Optional[torch.Tensor] = None):
        """Tensor-core prefill of one request into ``slot``; its first token comes out of the next ``batch_decode`` step."""
        p = np.ascontiguousarray(np.asarray(prompt.cpu() if isinstance(prompt, torch.Tensor) else prompt, dtype=np.int32))
        assert p.ndim == 2 and p.shape[0] == self.rows, f"prompt must be ({self.rows}, T)"
        if noise is not None:
            assert noise.dtype == torch.bfloat16 and noise.is_cuda and noise.is_contiguous()
        self._batch_noise[slot] = noise
        capi.check(self.lib.dualar_batch_prefill(self._h, int(slot), p.ctypes.data, p.shape[1], int(max_new_tokens),
                                                 float(temperature), float(top_p), float(repetition_penalty), int(seed),
                                                 _ptr(noise), self._stream()))

    def batch_decode(self, n_steps: int):
        capi.check(self.lib.dualar_batch_decode(self._h, int(n_steps), self._stream()))

    def batch_collect(self, slot: int, capacity: Optional[int] = None):
        """-> (tokens (C+1, n) int32 ndarray, finished) of one slot."""
        cap = capacity or self.cfg.max_seq_len
        out = np.zeros((self.rows, cap), dtype=np.int32)
        n, fin = C.c_int(0), C.c_int(0)
        capi.check(self.lib.dualar_batch_collect(self._h, int(slot), out.ctypes.data, cap, C.byref(n), C.byref(fin), self._stream()))
        return out[:, : n.value].copy(), bool(fin.value)

    def batch_release(self, slot: int):
        capi.check(self.lib.dualar_batch_release(self._h, int(slot)))

    def batch_read(self, name: str) -> torch.Tensor:
        cfg, B = self.cfg, self.max_batch
        shapes = {
            "slow_logits": ((B, cfg.vocab_size), torch.bfloat16), "slow_logits_raw": ((B, cfg.vocab_size), torch.bfloat16),
            "hidden": ((B, cfg.dim), torch.bfloat16), "fast_logits": ((B, cfg.num_codebooks - 1, self.fast_vocab), torch.bfloat16),
            "tokens": ((B, self.rows), torch.int32), "positions": ((B,), torch.int32), "done": ((B,), torch.int32),
            "n_gen": ((B,), torch.int32), "launches": ((1,), torch.int32), "bstep_phases": ((1,), torch.int32), "groups": ((1,), torch.int32),
        }
        if name in ("bstep_kinds", "bstep_timeline"):      # profiling of the persistent step (DUALAR_BS_TIMELINE=1): one entry per phase
            n = int(self.batch_read("bstep_phases")[0])
            shapes[name] = ((n,), torch.int32) if name == "bstep_kinds" else ((n, 2), torch.int64)
        shape, dt = shapes[name]
        out = torch.empty(shape, dtype=dt)
        capi.check(self.lib.dualar_batch_read(self._h, name.encode(), out.data_ptr(), out.numel() * out.element_size(), self._stream()))
        return out

    def set_option(self, name: str, value: float):
        capi.check(self.lib.dualar_set_option(self._h, name.encode(), float(value)))

    # ---- noise --------------------------------------------------------------------------------------
    def seed(self, seed: int):
        capi.check(self.lib.dualar_seed(self._h, int(seed)))

    def fill_noise(self, seed: int, step: int, head: int, n: int) -> torch.Tensor:
        out = torch.empty(n, dtype=torch.bfloat16, device=self.device)
        capi.check(self.lib.dualar_fill_noise(self._h, int(seed), int(step), int(head), out.data_ptr(), n, self._stream()))
        return out

    def step_noise(self, seed: int, step: int) -> torch.Tensor:
        """One step's noise block in the layout dualar_step expects."""
        parts = [self.fill_noise(seed, step, 0, self.cfg.vocab_size)]
        for k in range(1, self.cfg.num_codebooks):
            parts.append(self.fill_noise(seed, step, k, self.fast_vocab))
        return torch.cat(parts)

    def set_noise(self, noise: Optional[torch.Tensor]):
        """Explicit noise for the loop API: (n_steps * noise_per_step) bf16 on the device, or None."""
        self._noise_keepalive = noise
        if noise is None:
            capi.check(self.lib.dualar_set_noise(self._h, None, 0))
        else:
            assert noise.dtype == torch.bfloat16 and noise.is_cuda and noise.is_contiguous()
            capi.check(self.lib.dualar_set_noise(self._h, noise.data_ptr(), noise.numel() // self.noise_per_step))

    def debug_sample(self, head: int, logits: torch.Tensor, previous_tokens: Optional[torch.Tensor],
                     temperature: torch.Tensor, top_p: torch.Tensor, repetition_penalty: torch.Tensor,
                     noise: Optional[torch.Tensor]) -> int:
        """Test hook: the sampler of head ``head`` alone, on caller-supplied bf16 logits."""
        n = self.cfg.vocab_size if head == 0 else self.fast_vocab
        lg = logits.to(device=self.device, dtype=torch.bfloat16).contiguous()
        assert lg.numel() == n
        prev_ptr, prev_stride = None, 0
        if previous_tokens is not None:
            pt = previous_tokens.to(device=self.device, dtype=torch.int32).contiguous()
            assert pt.shape == (self.rows, 16)
            prev_ptr, prev_stride = pt.data_ptr(), pt.stride(0)
        out = torch.zeros(1, dtype=torch.int32, device=self.device)
        capi.check(self.lib.dualar_debug_sample(self._h, int(head), lg.data_ptr(), prev_ptr, prev_stride,
                                                temperature.data_ptr(), top_p.data_ptr(), repetition_penalty.data_ptr(),
                                                _ptr(noise), out.data_ptr(), self._stream()))
        return int(out.item())

    # ---- introspection --------------------------------------------------------------------------------
    def read(self, name: str) -> torch.Tensor:
        cfg = self.cfg
        shapes = {
            "slow_logits": ((cfg.vocab_size,), torch.bfloat16), "slow_logits_raw": ((cfg.vocab_size,), torch.bfloat16),
            "hidden": ((cfg.dim,), torch.bfloat16),
            "fast_logits": ((cfg.num_codebooks - 1, self.fast_vocab), torch.bfloat16),
            "tokens": ((self.rows,), torch.int32), "n_cand": ((1,), torch.int32), "prefix_reused": ((1,), torch.int32), "prefill_launches": ((1,), torch.int32), "nucleus": ((cfg.num_codebooks,), torch.int32),
            "qkv": (((cfg.n_head + 2 * cfg.n_local_heads) * cfg.head_dim,), torch.bfloat16),
            "y": ((cfg.n_head * cfg.head_dim,), torch.bfloat16), "h": ((cfg.dim,), torch.bfloat16),
            "act": ((cfg.intermediate_size,), torch.bfloat16), "fast_x": ((cfg.fast_dim,), torch.bfloat16),
            "fast_in": ((cfg.fast_dim,), torch.bfloat16), "timeline": ((2048, 8), torch.int64), "cand": ((8192,), torch.int64), "timeline2": ((400, 160, 4), torch.int64),
        }
        shape, dt = shapes[name]
        out = torch.empty(shape, dtype=dt)
        capi.check(self.lib.dualar_read_buffer(self._h, name.encode(), out.data_ptr(), out.numel() * out.element_size(),
                                               self._stream()))
        return out

    def launches_per_step(self) -> tuple[int, int]:
        a, b = C.c_int(0), C.c_int(0)
        capi.check(self.lib.dualar_launches_per_step(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def weight_bytes(self) -> tuple[int, int]:
        a, b = C.c_int64(0), C.c_int64(0)
        capi.check(self.lib.dualar_weight_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value
