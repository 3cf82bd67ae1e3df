"""Micro-benchmark of the fused attention kernels (sct_attn_fwd / sct_attn_bwd) on the cfg3 shapes, next to the
kernels they replace: torch.nn.functional.scaled_dot_product_attention (torch nn/functional.py:6244-6691, what
nn.MultiheadAttention calls) with the cuDNN, flash and memory-efficient back-ends on the same problems
(B=32, H=8, L=1024, dh=96, bf16; +-dropout, causal, key padding).
    python tools/attn_bench.py [--once] [--no-sdpa]     (--once: a single fwd+bwd per case, for ncu)"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sct_gan_b200 import kernels as kn  # noqa: E402

B, H, DH, L = 32, 8, 96, 1024
D = H * DH
REPS = 10


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS


def sdpa_case(backend, q, k, v, d_o, mask, causal, p):
    """fwd and fwd+bwd time of SDPA on [B, H, L, dh] tensors; returns (fwd ms, bwd ms) or an error string."""
    from torch.nn.attention import SDPBackend, sdpa_kernel

    be = {"cudnn": SDPBackend.CUDNN_ATTENTION, "flash": SDPBackend.FLASH_ATTENTION,
          "efficient": SDPBackend.EFFICIENT_ATTENTION}[backend]
    try:
        with sdpa_kernel([be]):
            def fwd():
                return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=p, is_causal=causal)

            t_f = timed(lambda: fwd())
            qg, kg, vg = (t.detach().requires_grad_(True) for t in (q, k, v))

            def fb():
                o = F.scaled_dot_product_attention(qg, kg, vg, attn_mask=mask, dropout_p=p, is_causal=causal)
                o.backward(d_o)
                qg.grad = kg.grad = vg.grad = None

            t_fb = timed(fb)
        return t_f, t_fb - t_f
    except Exception as e:  # back-end does not support the case (e.g. flash + arbitrary mask)
        return f"unsupported ({str(e).splitlines()[0][:60]})"


def main():
    once = "--once" in sys.argv
    sdpa = "--no-sdpa" not in sys.argv and not once
    torch.manual_seed(0)
    qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
    d_o = torch.randn(B * L, D, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    lens = torch.randint(L // 2, L + 1, (B,), device="cuda")
    kpm = (torch.arange(L, device="cuda")[None, :] >= lens[:, None]).contiguous()
    cases = [("self/no-mask p=0", None, False, 0.0), ("self/no-mask p=0.3", None, False, 0.3),
             ("self/kpm p=0.3", kpm, False, 0.3), ("causal p=0.3", None, True, 0.3)]
    # [B, H, L, dh] copies for SDPA (its preferred layout; the transposes are not timed)
    q4, k4, v4 = (qkv[:, i * D:(i + 1) * D].reshape(B, L, H, DH).transpose(1, 2).contiguous() for i in range(3))
    do4 = d_o.reshape(B, L, H, DH).transpose(1, 2).contiguous()
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
        res = [timed(fwd), timed(bwd)]
        fl = 4.0 * B * H * L * L * DH * (0.5 if causal else 1.0) / 1e12
        print(f"{name:20s} sct    fwd {res[0]*1e3:7.1f} us {fl / res[0] * 1e3:6.1f} TF/s   bwd {res[1]*1e3:7.1f} us "
              f"{2.5 * fl / res[1] * 1e3:6.1f} TF/s (algorithmic flops)", flush=True)
        if sdpa:
            amask = None if mask is None else (~mask)[:, None, None, :].expand(B, 1, L, L)  # True = attend
            for be in ("cudnn", "flash", "efficient"):
                r = sdpa_case(be, q4, k4, v4, do4, amask, causal, p)
                if isinstance(r, str):
                    print(f"{'':20s} {be:9s} {r}", flush=True)
                else:
                    print(f"{'':20s} {be:9s} fwd {r[0]*1e3:7.1f} us {fl / r[0] * 1e3:6.1f} TF/s   bwd {r[1]*1e3:7.1f} us "
                          f"{2.5 * fl / r[1] * 1e3:6.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
