"""Micro-benchmark of the fused attention kernels (sct_attn_fwd / sct_attn_bwd) on the cfg3 shapes.
    python tools/attn_bench.py [--once]     (--once: a single fwd+bwd per case, for ncu)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sct_gan_b200 import kernels as kn  # noqa: E402

B, H, DH, L = 32, 8, 96, 1024
D = H * DH


def main():
    once = "--once" in sys.argv
    torch.manual_seed(0)
    qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
    d_o = torch.randn(B * L, D, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    lens = torch.randint(L // 2, L + 1, (B,), device="cuda")
    kpm = (torch.arange(L, device="cuda")[None, :] >= lens[:, None]).contiguous()
    cases = [("self/no-mask p=0", None, False, 0.0), ("self/no-mask p=0.3", None, False, 0.3),
             ("self/kpm p=0.3", kpm, False, 0.3), ("causal p=0.3", None, True, 0.3)]
    for name, mask, causal, p in cases:
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]

        def fwd():
            return kn.attn_fwd(q, k, v, B, H, L, L, kpm=mask, causal=causal, p_drop=p, seed=1, offset=2)

        o, lse = fwd()

        def bwd():
            kn.attn_bwd(q, k, v, o, d_o, lse, B, H, L, L, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], kpm=mask,
                        causal=causal, p_drop=p, seed=1, offset=2)

        bwd()
        torch.cuda.synchronize()
        if once:
            continue
        res = []
        for fn in (fwd, bwd):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 10)
        fl = 4.0 * B * H * L * L * DH * (0.5 if causal else 1.0) / 1e12
        print(f"{name:20s} fwd {res[0]*1e3:7.1f} us {fl / res[0] * 1e3:6.1f} TF/s   bwd {res[1]*1e3:7.1f} us "
              f"{2.5 * fl / res[1] * 1e3:6.1f} TF/s (algorithmic flops)", flush=True)


if __name__ == "__main__":
    main()
