"""Micro-benchmark of the feed-forward GEMMs with fused epilogues against the plain GEMMs of the same shape
(M = 32768 rows, d = 768, ff = 2048): linear1 (+ GELU + dropout), linear2's dgrad (* local derivative).
    python tools/ffn_bench.py          (SCT_EPI_DBG=1|2|4 switches parts of the epilogue off for timing experiments)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sct_gan_b200 import kernels as kn  # noqa: E402

M, D, FF = 32768, 768, 2048


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
    x = torch.randn(M, D, device="cuda").bfloat16()
    w1 = (torch.randn(FF, D, device="cuda") * 0.05).bfloat16()
    b1 = torch.randn(FF, device="cuda")
    w2 = (torch.randn(D, FF, device="cuda") * 0.05).bfloat16()
    dy = torch.randn(M, D, device="cuda").bfloat16()
    z = torch.empty(M, FF, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * FF * D / 1e12
    for name, fn in (("linear1 plain nt", lambda: kn.gemm_nt(x, w1, b1, out=z)),
                     ("linear1 + gelu + dropout(0.3)", lambda: kn.gemm_nt_gelu(x, w1, b1, 0.3, 1, 2)),
                     ("linear1 + gelu, p = 0", lambda: kn.gemm_nt_gelu(x, w1, b1, 0.0, 1, 2)),
                     ("linear2 dgrad plain nn", lambda: kn.gemm_nn(dy, w2, out=z)),
                     ("linear2 dgrad * G", lambda: kn.gemm_nn_mul(dy, w2, z)),
                     ("gelu_dropout fwd (separate)", lambda: kn.gelu_dropout_fwd(z, 0.3, 1, 2)),
                     ("gelu_dropout bwd (separate)", lambda: kn.gelu_dropout_bwd(z, z, 0.3, 1, 2))):
        t = timeit(fn)
        print(f"{name:34s} {t * 1e3:7.1f} us  {fl / t * 1e3:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
