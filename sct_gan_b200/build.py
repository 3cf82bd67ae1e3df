"""Builds libsct_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Cross-compiles without a GPU.  Objects go to sct_gan_b200/csrc/build/, the library to
sct_gan_b200/libsct_b200.so (git-ignored, shipped to the GPU box with the working tree).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = CSRC / "build"
LIB_PATH = PKG_DIR / "libsct_b200.so"
INCLUDE = PKG_DIR.parent / "include"

SOURCES = ["api.cu", "gemm.cu", "attn.cu", "rowwise.cu", "loss.cu", "optim.cu", "sample.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
# --use_fast_math (approximate division / transcendentals, flush-to-zero) for the activation-side kernels only: the
# optimiser (AdamW bias corrections, denormal-range second moments with eps = 1e-9) and the loss kernels keep IEEE
# arithmetic so the update matches torch.optim.AdamW
FAST_MATH = {"gemm.cu", "attn.cu", "rowwise.cu", "api.cu"}


NVCC_FLAGS += os.environ.get("SCT_NVCC_EXTRA", "").split()  # e.g. -DSCT_ATTN_TRACE for instrumented builds


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build libsct_b200.so")
    return cand


def _newer(a: Path, deps: list[Path]) -> bool:
    if not a.exists():
        return False
    t = a.stat().st_mtime
    return all(d.stat().st_mtime <= t for d in deps if d.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    headers = [CSRC / "common.cuh", INCLUDE / "sct_b200.h", Path(__file__)]
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    objs = [BUILD_DIR / (s.stem + ".o") for s in srcs]

    def compile_one(pair):
        src, obj = pair
        if not force and _newer(obj, [src] + headers):
            return None
        cmd = [nvcc, *NVCC_FLAGS, *(["--use_fast_math"] if src.name in FAST_MATH else []), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return r.stderr if verbose else None

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        logs = list(ex.map(compile_one, zip(srcs, objs)))
    if verbose:
        for log in logs:
            if log:
                print(log, file=sys.stderr)
    if force or not _newer(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs),
               "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
