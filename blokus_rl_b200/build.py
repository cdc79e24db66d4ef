"""In-tree build of libblokus_b200.so (nvcc, sm_100a only)."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
SRC = PKG / "csrc" / "blk_engine.cu"
DEPS = [SRC, PKG / "csrc" / "blk_orient.inc", ROOT / "include" / "blokus_b200.h"]
OUT = PKG / "libblokus_b200.so"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in DEPS):
        return OUT
    cmd = ["nvcc", *NVCC_FLAGS, *( ["-Xptxas", "-v"] if verbose else []), "-o", str(OUT), str(SRC)]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
