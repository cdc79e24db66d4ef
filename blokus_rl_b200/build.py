"""In-tree build of libblokus_b200.so (nvcc, sm_100a only).

One object per kernel specialisation (blk_inst.cu compiled with -DBLK_INST_N/-DBLK_INST_P) plus the host /
C-ABI object, compiled in parallel and linked into blokus_rl_b200/libblokus_b200.so.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
OUT = PKG / "libblokus_b200.so"
DEPS = [CSRC / "blk_engine.cu", CSRC / "blk_inst.cu", CSRC / "blk_puct.cu", CSRC / "blk_kernels.cuh", CSRC / "blk_search.cuh", CSRC / "blk_orient.inc",
        CSRC / "blk_small.cu", CSRC / "blk_small_fields.inc",
        ROOT / "include" / "blokus_b200.h"]
GEOMETRIES = [(20, 4), (20, 2), (14, 4), (14, 2), (7, 2), (0, 0)]
SMALL_SIZES = [5, 6, 7]          # thread-per-env kernels on 64-bit bitboards (csrc/blk_small.cu)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def kernel_source_hash() -> str:
    """sha256 over the CUDA sources and the ABI header: identifies the kernels a profile / traffic figure was taken on."""
    import hashlib
    h = hashlib.sha256()
    for d in sorted(DEPS, key=lambda p: p.name):
        h.update(d.name.encode())
        h.update(d.read_bytes())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None, out: Path | None = None) -> Path:
    """``out``: build a VARIANT library there (own object directory next to it) -- A/B runs of the compile-time knobs
    (BLK_WARPS, BLK_ROLL_WARPS, ...) select it at run time with BLOKUS_B200_LIB; the product library is untouched."""
    global OUT, OBJ
    if out is not None:
        saved = (OUT, OBJ)
        OUT, OBJ = Path(out), Path(out).with_suffix(".objs")
        try:
            return build(True, verbose, extra_flags)
        finally:
            OUT, OBJ = saved
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in DEPS):
        return OUT
    OBJ.mkdir(exist_ok=True)
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (extra_flags or [])
    jobs = [(OBJ / "blk_engine.o", ["nvcc", *flags, "-c", str(CSRC / "blk_engine.cu")]),
            (OBJ / "blk_puct.o", ["nvcc", *flags, "-c", str(CSRC / "blk_puct.cu")])]
    for n, p in GEOMETRIES:
        jobs.append((OBJ / f"blk_inst_{n}_{p}.o",
                     ["nvcc", *flags, f"-DBLK_INST_N={n}", f"-DBLK_INST_P={p}", "-c", str(CSRC / "blk_inst.cu")]))

    for n in SMALL_SIZES:
        jobs.append((OBJ / f"blk_small_{n}.o", ["nvcc", *flags, f"-DBLK_SMALL_N={n}", "-c", str(CSRC / "blk_small.cu")]))

    def run(job):
        obj, cmd = job
        res = subprocess.run(cmd + ["-o", str(obj)], capture_output=True, text=True)
        if verbose:
            sys.stderr.write(res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {obj.name}:\n{res.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(run, jobs))
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(OUT), *map(str, objs)])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
