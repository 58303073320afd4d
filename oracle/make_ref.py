"""ORACLE SUPPORT (test infrastructure) -- make the UNMODIFIED reference travel to the GPU box.

    python oracle/make_ref.py        (also run by __graft_entry__.build() whenever /root/reference is present)

The reference is pure Python.  `/root/reference` exists only in the build container, so this recipe copies its package
(`fish_tts/`, nothing else) byte for byte into `oracle/_ref/` -- an OUTPUT directory that is git-ignored (no reference source
enters the history) but not gpurun-ignored, so it ships with the repo snapshot like a built `.so`.  `oracle/ref_harness.py`
imports the reference from `/root/reference` when that exists and from `oracle/_ref/` otherwise; `bench.py --impl reference`
and `cpu_baseline` / `torch_baselines` then time the reference's own `init_model` / `generate` (kind "reference") instead of
the oracle port.  The `dac` / `audiotools` stand-ins the import needs are installed in `sys.modules` by the harness, not
written here.  A manifest with the sha256 of every copied file is written next to the copy.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

SRC = Path("/root/reference")
DST = Path(__file__).resolve().parent / "_ref"


def make_ref(verbose: bool = True) -> bool:
    if not (SRC / "fish_tts" / "models" / "inference.py").exists():
        if verbose:
            print(f"[make_ref] {SRC} not present: keeping whatever is under {DST}")
        return DST.exists()
    if DST.exists():
        shutil.rmtree(DST)
    manifest = {}
    for p in sorted((SRC / "fish_tts").rglob("*.py")):
        rel = p.relative_to(SRC)
        out = DST / rel
        out.parent.mkdir(parents=True, exist_ok=True)
        data = p.read_bytes()
        out.write_bytes(data)
        manifest[str(rel)] = hashlib.sha256(data).hexdigest()
    (DST / "MANIFEST.json").write_text(json.dumps({"source": str(SRC), "files": manifest}, indent=1))
    if verbose:
        print(f"[make_ref] copied {len(manifest)} files of the unmodified reference to {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
