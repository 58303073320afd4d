"""ctypes binding of include/dualar.h -- the only way the Python host reaches the GPU.

There is deliberately no fallback: if libdualar.so is missing or a call fails this raises.
(INTEGRATION.md shows the same binding as a maintainer of the reference would add it.)
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libdualar.so"
ABI_VERSION = 1

# every symbol include/dualar.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "dualar_abi_version", "dualar_last_error", "dualar_create", "dualar_load_weight", "dualar_finalize",
    "dualar_destroy", "dualar_bind_kv", "dualar_step", "dualar_prefill", "dualar_decode", "dualar_collect",
    "dualar_generate", "dualar_seed", "dualar_fill_noise", "dualar_set_noise", "dualar_read_buffer",
    "dualar_launches_per_step", "dualar_weight_bytes", "dualar_debug_sample", "dualar_set_option",
    "dualar_batch_init", "dualar_batch_prefill", "dualar_batch_decode", "dualar_batch_collect", "dualar_batch_release",
    "dualar_batch_read", "dualar_decode_async", "dualar_wait",
]

_I32_FIELDS = [
    "abi_version", "vocab_size", "n_layer", "n_head", "dim", "intermediate_size", "n_local_heads", "head_dim",
    "max_seq_len", "codebook_size", "num_codebooks", "n_fast_layer", "fast_dim", "fast_n_head",
    "fast_n_local_heads", "fast_head_dim", "fast_intermediate_size", "tie_word_embeddings",
    "attention_qkv_bias", "attention_o_bias", "attention_qk_norm", "fast_attention_qkv_bias",
    "fast_attention_o_bias", "fast_attention_qk_norm", "scale_codebook_embeddings", "semantic_begin_id",
    "semantic_end_id", "im_end_id",
]


class DualarConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in _I32_FIELDS] + [("rope_base", C.c_float), ("norm_eps", C.c_float)]


class DualarError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libdualar error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libdualar.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m fish_tts_b200._build` "
            "(there is no CPU or PyTorch fallback for the decode path)")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)
    sig = {
        "dualar_abi_version": (C.c_int, []),
        "dualar_last_error": (C.c_char_p, []),
        "dualar_create": (C.c_int, [C.POINTER(DualarConfig), C.c_int, C.POINTER(vp)]),
        "dualar_load_weight": (C.c_int, [vp, C.c_char_p, vp, C.c_int64, C.c_int]),
        "dualar_finalize": (C.c_int, [vp]),
        "dualar_destroy": (None, [vp]),
        "dualar_bind_kv": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
        "dualar_step": (C.c_int, [vp, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp]),
        "dualar_prefill": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, vp]),
        "dualar_decode": (C.c_int, [vp, C.c_int, vp]),
        "dualar_collect": (C.c_int, [vp, vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), vp]),
        "dualar_generate": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, vp, C.c_int,
                                      C.POINTER(C.c_int), vp]),
        "dualar_seed": (C.c_int, [vp, C.c_uint64]),
        "dualar_fill_noise": (C.c_int, [vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, C.c_int64, vp]),
        "dualar_set_noise": (C.c_int, [vp, vp, C.c_int64]),
        "dualar_read_buffer": (C.c_int, [vp, C.c_char_p, vp, C.c_int64, vp]),
        "dualar_set_option": (C.c_int, [vp, C.c_char_p, C.c_double]),
        "dualar_debug_sample": (C.c_int, [vp, C.c_int, vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp]),
        "dualar_launches_per_step": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "dualar_weight_bytes": (C.c_int, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        "dualar_batch_init": (C.c_int, [vp, C.c_int, C.c_int]),
        "dualar_batch_prefill": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint64, vp, vp]),
        "dualar_batch_decode": (C.c_int, [vp, C.c_int, vp]),
        "dualar_batch_collect": (C.c_int, [vp, C.c_int, vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), vp]),
        "dualar_batch_release": (C.c_int, [vp, C.c_int]),
        "dualar_batch_read": (C.c_int, [vp, C.c_char_p, vp, C.c_int64, vp]),
        "dualar_decode_async": (C.c_int, [vp, C.c_int, vp, vp, vp, C.POINTER(C.c_int)]),
        "dualar_wait": (C.c_int, [vp, C.c_int, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.dualar_abi_version() != ABI_VERSION:
        raise ImportError(f"libdualar.so ABI {lib.dualar_abi_version()} != {ABI_VERSION}; rebuild it")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise DualarError(rc, load().dualar_last_error().decode())


def make_config(cfg) -> DualarConfig:
    """fish_tts_b200.config.DualARConfig -> the C struct."""
    c = DualarConfig()
    c.abi_version = ABI_VERSION
    for n in _I32_FIELDS[1:]:
        setattr(c, n, int(getattr(cfg, n)))
    c.rope_base, c.norm_eps = float(cfg.rope_base), float(cfg.norm_eps)
    return c
