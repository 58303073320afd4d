"""GPU-side helpers shared by the parity tests and the diagnostic script (oracle = checker only)."""
from __future__ import annotations

import torch
from torch.nn.attention import SDPBackend, sdpa_kernel

from fish_tts_b200.engine import DualAREngine
from fish_tts_b200.synthetic import make_state_dict
from oracle import dualar_oracle as orc


def build_pair(cfg, seed=0, bind_kv=True, options=None, sd=None, **sd_kw):
    """oracle on cuda:0 + engine; with bind_kv the engine runs on the oracle's KV tensors (phase A)."""
    sd = sd or make_state_dict(cfg, seed=seed, **sd_kw)
    m = orc.OracleModel.build(cfg, sd, device="cuda:0")
    m.setup_caches(cfg.max_seq_len)
    kv = None
    if bind_kv:
        kv = {"slow": [(k, v) for k, v in m.kv], "fast": [(k, v) for k, v in m.fast_kv]}
    eng = DualAREngine(cfg, sd, device=0, kv=kv, options=options)
    return m, eng, sd


def block_noise_source(cfg, block: torch.Tensor) -> orc.NoiseSource:
    """feed one step's engine-layout noise block to the oracle in its sampling order"""
    fv = min(1024, cfg.codebook_size)

    def fn(call, n):
        off = 0 if call == 0 else cfg.vocab_size + (call - 1) * fv
        return block[off: off + n]

    return orc.NoiseSource(fn)


class TeacherForced:
    """Runs oracle and engine side by side on the ORACLE's trajectory and KV state, one step at a time."""

    def __init__(self, cfg, m, eng, prompt, T, p, rp, noise_seed=11, alt=None):
        """alt: a second OracleModel (e.g. on the CPU) stepped along the same trajectory -- the yardstick for how far
        two of the reference's own backends are from each other."""
        self.cfg, self.m, self.eng, self.alt = cfg, m, eng, alt
        dev = m.device
        self.t = [torch.tensor(v, device=dev, dtype=torch.float) for v in (T, p, rp)]
        self.C1 = cfg.num_codebooks + 1
        self.noise_seed = noise_seed
        Tlen = prompt.size(1)
        self.prev = torch.zeros((self.C1, cfg.max_seq_len), dtype=torch.int32, device=dev)
        blk = eng.step_noise(noise_seed, 0)
        with torch.inference_mode(), sdpa_kernel(SDPBackend.MATH):
            first = orc.decode_one_token_ar(m, prompt.view(1, self.C1, -1).to(dev), torch.arange(Tlen, device=dev),
                                            *self.t, None, noise=block_noise_source(cfg, blk), stable_ties=True)
        if alt is not None:
            alt.setup_caches(cfg.max_seq_len)
            ta = [x.to(alt.device) for x in self.t]
            with torch.inference_mode(), sdpa_kernel(SDPBackend.MATH):
                orc.decode_one_token_ar(alt, prompt.view(1, self.C1, -1).to(alt.device), torch.arange(Tlen, device=alt.device),
                                        *ta, None, noise=block_noise_source(cfg, blk.to(alt.device)), stable_ties=True)
        self.cur = first.view(1, self.C1, -1).clone()
        self.input_pos = torch.tensor([Tlen], device=dev, dtype=torch.int32)
        self.i = 0

    def step(self):
        """-> dict(mine tokens, ref tokens, my/ref logits...)"""
        cfg, m, eng, i = self.cfg, self.m, self.eng, self.i
        window = self.prev[:, :16] if i < 16 else self.prev[:, i - 16: i]
        blk = eng.step_noise(self.noise_seed, i + 1)
        mine = eng.step(self.cur, self.input_pos, window, *self.t, noise=blk).clone()
        torch.cuda.synchronize()
        out = {"mine": mine[:, 0].cpu(), "my_slow": eng.read("slow_logits_raw"), "my_fast": eng.read("fast_logits"),
               "my_hidden": eng.read("hidden"), "nucleus": eng.read("nucleus")}
        tr = []
        with torch.inference_mode(), sdpa_kernel(SDPBackend.MATH):
            ref = orc.decode_one_token_ar(m, self.cur, self.input_pos, *self.t, window,
                                          noise=block_noise_source(cfg, blk), stable_ties=True, trace=tr)
        out.update(ref=ref[:, 0].cpu(), ref_slow=tr[0].slow_logits.cpu(), ref_fast=torch.stack(tr[0].fast_logits).cpu(),
                   ref_hidden=tr[0].hidden.cpu(), window=window.clone().cpu(), noise=blk)
        if self.alt is not None:
            alt, tra = self.alt, []
            ta = [x.to(alt.device) for x in self.t]
            with torch.inference_mode(), sdpa_kernel(SDPBackend.MATH):
                ra = orc.decode_one_token_ar(alt, self.cur.to(alt.device), self.input_pos.to(alt.device), *ta, window.to(alt.device),
                                             noise=block_noise_source(cfg, blk.to(alt.device)), stable_ties=True, trace=tra)
            out.update(alt=ra[:, 0].cpu(), alt_slow=tra[0].slow_logits.cpu(), alt_fast=torch.stack(tra[0].fast_logits).cpu(),
                       alt_hidden=tra[0].hidden.cpu())
        self.input_pos += 1
        self.cur = ref.view(1, self.C1, -1).clone()
        self.prev[:, i: i + 1] = ref
        self.i += 1
        return out


def oracle_generate_tokenwise_prefill(m, prompt, n_new, T, p, rp, noise_fn):
    """`generate` of the oracle, except that the prompt is pushed through the model one position at a time -- the
    engine's round-1 prefill does the same (every GEMM is then the M=1 case on both sides), so the two KV states
    agree and the test isolates the loop machinery (window, position, EOS, noise bookkeeping)."""
    cfg, dev = m.cfg, m.device
    C1 = cfg.num_codebooks + 1
    m.setup_caches(cfg.max_seq_len)
    t = [torch.tensor(v, device=dev, dtype=torch.float) for v in (T, p, rp)]
    prompt = prompt.to(dev)
    Tlen = prompt.size(1)
    noise = orc.NoiseSource(noise_fn)
    with torch.inference_mode(), sdpa_kernel(SDPBackend.MATH):
        for i in range(Tlen - 1):
            orc.forward_generate(m, prompt[:, i:i + 1].view(1, C1, 1), torch.tensor([i], device=dev))
        first = orc.decode_one_token_ar(m, prompt[:, -1:].view(1, C1, 1), torch.tensor([Tlen - 1], device=dev), *t, None,
                                        noise=noise, stable_ties=True)
        rest = orc.decode_n_tokens(m, first.view(1, C1, -1), torch.tensor([Tlen], device=dev, dtype=torch.int), n_new - 1, *t,
                                   noise=noise, stable_ties=True)
    return torch.cat([first, rest], dim=1)
