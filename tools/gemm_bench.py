"""Micro-benchmark of the tcgen05 GEMM (sct_gemm_bf16_{nt,nn,tn}) on the shapes of the cfg3 step.
    python tools/gemm_bench.py            (needs a B200; prints TFLOP/s per shape / variant / N-tile)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sct_gan_b200 import kernels as kn  # noqa: E402

M = 32768
SHAPES = [(2304, 768), (768, 768), (1536, 768), (2048, 768), (768, 2048), (768, 1536), (384, 768), (768, 384)]


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for N, K in SHAPES:
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        dy = torch.randn(M, N, device="cuda").bfloat16()
        bias = torch.randn(N, device="cuda")
        dw = torch.zeros(N, K, device="cuda")
        fl = 2.0 * M * N * K / 1e12
        row = [f"N={N:5d} K={K:5d}"]
        for bn in (128, 256):
            y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            t = timeit(lambda: kn.gemm_nt(x, w, bias, out=y, bn=bn))
            row.append(f"nt{bn}: {fl / t * 1e3:7.1f}")
        for bn in (128, 256):
            dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
            t = timeit(lambda: kn.gemm_nn(dy, w, out=dx, bn=bn))
            row.append(f"nn{bn}: {fl / t * 1e3:7.1f}")
        for ks in (0, 4, 16):
            t = timeit(lambda: kn.gemm_tn(dy, x, dw, k_splits=ks))
            row.append(f"tn/ks{ks}: {fl / t * 1e3:7.1f}")
        t = timeit(lambda: torch.matmul(x, w.t()))
        row.append(f"cublas nt: {fl / t * 1e3:7.1f}")
        print("  ".join(row), flush=True)
    del flush


if __name__ == "__main__":
    main()
