"""Micro-benchmark of the HBM-bound row kernels at the cfg3 per-GPU shapes (rows = 32768, d = 768, ff = 2048).
    python tools/rowwise_bench.py        (needs a B200; prints us per launch and achieved GB/s of algorithmic bytes)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sct_gan_b200 import kernels as kn  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32


def timeit(fn, flush, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()  # inputs are evicted from the 126 MB L2 between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


def main():
    torch.manual_seed(0)
    dev = "cuda"
    peak = 6449.4
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            peak = float(json.load(open(p)).get("hbm_copy_GBps", peak))
        except Exception:
            pass
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    R, d, ff, V = 32768, 768, 2048, 50272
    x = torch.randn(R, d, device=dev)
    br = torch.randn(R, d, device=dev).to(BF16)
    gam, bet = torch.randn(d, device=dev), torch.randn(d, device=dev)
    z = torch.randn(R, ff, device=dev).to(BF16)
    gh = torch.randn(R, ff, device=dev).to(BF16)
    rows = []

    def rec(name, us, nbytes):
        rows.append((name, us, nbytes / us / 1e3))

    us = timeit(lambda: kn.add_dropout_ln_fwd(x, br, 1.0, gam, bet, p_drop=0.3, seed=1, offset=1), flush)
    rec("add_dropout_ln_fwd (x, branch -> x', LN)", us, R * d * (4 + 2 + 4 + 2))
    xo, yl, _, st = kn.add_dropout_ln_fwd(x, br, 1.0, gam, bet, p_drop=0.3, seed=1, offset=1)
    gx, gy = torch.randn(R, d, device=dev), torch.randn(R, d, device=dev).to(BF16)
    dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    us = timeit(lambda: kn.add_dropout_ln_bwd(gx, gy, None, xo, st, gam, 1.0, dg, db, p_drop=0.3, seed=1, offset=1), flush)
    rec("add_dropout_ln_bwd (g_x, g_ln, x' -> g_x, g_branch)", us, R * d * (4 + 2 + 4 + 4 + 2))
    us = timeit(lambda: kn.gelu_dropout_fwd(z, p_drop=0.3, seed=1, offset=2), flush)
    rec("gelu_dropout_fwd", us, R * ff * 4)
    us = timeit(lambda: kn.gelu_dropout_bwd(gh, z, p_drop=0.3, seed=1, offset=2), flush)
    rec("gelu_dropout_bwd", us, R * ff * 6)
    lg = torch.randn(4096, V, device=dev).to(BF16)
    tg = torch.randint(0, 50265, (4096,), device=dev)
    us = timeit(lambda: kn.ce_rows(lg, tg, 50265, grad_scale=1e-4, write_grad=True), flush)
    rec("ce_rows (4096 x 50265, loss + in-place gradient)", us, 4096 * V * 6)
    out = torch.zeros(ff, device=dev)
    us = timeit(lambda: kn.colsum_bf16(z, out), flush)
    rec("colsum_bf16 (32768 x 2048)", us, R * ff * 2)
    dst = torch.empty(R, 2 * d, device=dev, dtype=BF16)
    us = timeit(lambda: kn.cast_scale(x, dst, col_off=0, scale=1.0), flush)
    rec("cast_scale fp32 -> bf16", us, R * d * 6)
    for name, us, gbs in rows:
        print(f"{name:55s} {us:8.1f} us  {gbs:7.1f} GB/s  {gbs / peak:5.2f} of {peak:.0f}", flush=True)


if __name__ == "__main__":
    main()
