"""Build libdualar.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

`python -m fish_tts_b200._build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libdualar.so"
SOURCES = [CSRC / "engine.cu"]
HEADERS = sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "dualar.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # Blackwell B200 only; no PTX fallback for other archs
    "-lineinfo", "-O3", "-std=c++17", *os.environ.get("DUALAR_NVCC_DEFS", "").split(),      # experiment switches, e.g. DUALAR_NVCC_DEFS="-DDA_TC_MIN_BLOCKS=3"
    "-shared", "-Xcompiler", "-fPIC",
]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB), *map(str, SOURCES)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed ({r.returncode}):\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
