"""Host-side logic that needs no GPU: the gradient arena's hand-off to autograd, the bench's config / clock plumbing."""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _TwoUse(torch.autograd.Function):
    """A parameter used at two sites of one step: both backward passes accumulate into ONE arena buffer, the first
    hands it to autograd, the second returns None (the pattern of ops.EmbedLnPe)."""

    @staticmethod
    def forward(ctx, w, scale):
        ctx.save_for_backward(w)
        ctx.scale = scale
        return w * scale

    @staticmethod
    def backward(ctx, g):
        from sct_gan_b200 import ops

        (w,) = ctx.saved_tensors
        buf, first = ops.ARENA.zeros_for(w, w.shape, w.device)
        buf.add_(g * ctx.scale)
        return (buf if first else None), None


def test_grad_arena_buffers_are_adopted_not_cloned():
    from sct_gan_b200 import ops

    arena = ops.ARENA
    saved = (arena.buf, arena.need_last, arena._old)
    arena.buf, arena.need_last, arena._old = None, 0, []
    try:
        w = torch.randn(300, 7, requires_grad=True)
        dev = w.device

        def step():
            w.grad = None
            arena.begin(dev)
            (_TwoUse.apply(w, 2.0).sum() + _TwoUse.apply(w, 3.0).square().sum()).backward()
            arena.end()

        step()  # sizing pass: plain zeros
        want = w.grad.clone()
        step()  # arena pass
        lo = arena.buf.data_ptr()
        hi = lo + arena.buf.numel() * 4
        assert lo <= w.grad.data_ptr() < hi          # autograd adopted the arena view: no clone of the gradient
        assert torch.allclose(w.grad, want)
        ptr = w.grad.data_ptr()
        step()
        assert w.grad.data_ptr() == ptr              # same address every step (optimiser pointer tables stay valid)
        assert torch.allclose(w.grad, want)          # ... and the arena was cleared in between
    finally:
        arena.buf, arena.need_last, arena._old = saved
        arena.active = False


def _bench():
    spec = importlib.util.spec_from_file_location("sct_bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sct_bench_module"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_bench_config_is_shared_by_both_arms_and_names_the_workload():
    b = _bench()

    class A:
        no_vuln_heads, no_graph = False, False

    for name in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
        b.CFG.clear()
        b.CFG.update(b.CONFIGS[name], name=name)
        c1, c8 = b.config_record(1, A), b.config_record(8, A)
        assert name in c1["workload"] and "model" not in c1
        assert c8["global_batch"] == 8 * c1["global_batch"] and c1["seq_len"] == b.CFG["S"]
        assert c1["parallelism"].startswith("dp1") and c8["parallelism"].startswith("dp8")
    assert b.metric_name() == "generation_new_tokens_per_sec"  # cfg5 is the decode leg


def test_clock_sampler_survives_a_missing_nvidia_smi(monkeypatch):
    b = _bench()
    monkeypatch.setenv("PATH", "/nonexistent")
    s = b.ClockSampler(0).start()
    assert s.ready()
    with s as k:
        pass
    out = k.summary()
    assert out["samples"] == 0 and out["sm_mhz"] is None and out["reasons"] == []
