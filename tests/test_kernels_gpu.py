"""Kernel-level parity: every C-ABI entry point against a plain PyTorch fp32 reference of the same op.

Tolerances (stated per test): operands/outputs are bf16 with fp32 accumulation, so elementwise
agreement is a few bf16 ulps of the output scale; integer work (gather rows, masks) is bit-exact.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16, F32 = torch.bfloat16, torch.float32


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def _no_timeouts():
    from sct_gan_b200 import _lib

    torch.cuda.synchronize()
    assert _lib.load().sct_debug_timeouts() == 0, "a bounded mbarrier wait timed out"


# ------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 64, 128), (256, 256, 128, 128), (4096, 768, 768, 128), (4096, 2304, 768, 256),
    (4096, 2048, 768, 256), (4096, 768, 2048, 128), (1000, 520, 776, 128), (2048, 50265, 768, 256),
    (333, 384, 768, 128), (4096, 768, 384, 128), (1000, 520, 776, 256), (384, 512, 128, 256), (129, 768, 768, 256),
    (128, 768, 768, 64), (128, 2304, 768, 64), (100, 50265, 768, 64), (128, 768, 2048, 64), (300, 200, 72, 64),
])
def test_gemm_nt(cuda_dev, M, N, K, bn):
    from sct_gan_b200 import kernels as kn

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(BF16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(BF16)
    bias = torch.randn(N, device="cuda", generator=g)
    ldd = (N + 7) // 8 * 8
    buf = torch.zeros(M, ldd, device="cuda", dtype=BF16)
    out = kn.gemm_nt(a, w, bias, out=buf[:, :N], bn=bn)
    _no_timeouts()
    ref = a.float() @ w.float().t() + bias
    # bf16 output rounding (2^-9 relative) dominates; fp32 accumulation order differs from torch
    assert rel_l2(out, ref) < 4e-3
    assert (out.float() - ref).abs().max().item() < 0.05 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 64, 128), (4096, 768, 2304, 128), (4096, 768, 768, 128), (4096, 2048, 768, 256),
    (1000, 776, 520, 128), (2048, 768, 50272, 128), (1000, 776, 520, 256), (4096, 768, 2304, 256),
])
def test_gemm_nn(cuda_dev, M, N, K, bn):
    from sct_gan_b200 import kernels as kn

    g = torch.Generator(device="cuda").manual_seed(M + N + K + 1)
    a = (torch.randn(M, K, device="cuda", generator=g) * (1.0 / math.sqrt(K))).to(BF16)
    w = torch.randn(K, N, device="cuda", generator=g).to(BF16)
    out = kn.gemm_nn(a, w, bn=bn)
    _no_timeouts()
    ref = a.float() @ w.float()
    assert rel_l2(out, ref) < 4e-3


@pytest.mark.parametrize("M,N,K,ks", [
    (128, 128, 64, 1), (768, 768, 4096, 0), (2304, 768, 4096, 0), (768, 2048, 4096, 4),
    (520, 776, 1000, 0), (50265, 768, 2048, 1),
])
def test_gemm_tn(cuda_dev, M, N, K, ks):
    from sct_gan_b200 import kernels as kn

    g = torch.Generator(device="cuda").manual_seed(M + N + K + 2)
    lda = (M + 7) // 8 * 8  # ragged M (the vocab) lives in a pitch-padded scratch, as in VocabCE
    a = (torch.randn(K, lda, device="cuda", generator=g) * (1.0 / math.sqrt(K))).to(BF16)[:, :M]
    b = torch.randn(K, N, device="cuda", generator=g).to(BF16)
    base = torch.randn(M, N, device="cuda", generator=g)
    out = base.clone()
    kn.gemm_tn(a, b, out, alpha=0.5, k_splits=ks)
    _no_timeouts()
    ref = base + 0.5 * (a.float().t() @ b.float())
    # fp32 output: only accumulation-order differences
    assert rel_l2(out, ref) < 1e-4
    # same call with the fused bias gradient: colsum[m] += alpha * sum_k a[k, m], D unchanged by it
    out2 = base.clone()
    cs0 = torch.randn(M, device="cuda", generator=g)
    cs = cs0.clone()
    kn.gemm_tn(a, b, out2, alpha=0.5, k_splits=ks, colsum=cs)
    _no_timeouts()
    assert rel_l2(out2, ref) < 1e-4
    cs_ref = cs0 + 0.5 * a.float().sum(0)
    assert (cs - cs_ref).abs().max().item() < 1e-3 * max(1.0, cs_ref.abs().max().item())


# -------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, kpm, causal, scale):
    # q [B,H,Lq,dh] fp32 ...
    s = torch.einsum("bhqd,bhkd->bhqk", q, k) * scale
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :].bool(), float("-inf"))
    if causal:
        Lq, Lk = s.shape[-2:]
        s = s.masked_fill(torch.ones(Lq, Lk, device=s.device, dtype=torch.bool).triu(1), float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.einsum("bhqk,bhkd->bhqd", p, v)


@pytest.mark.parametrize("B,Lq,Lk,causal,masked", [
    (2, 128, 128, False, False), (2, 256, 256, True, False), (2, 512, 512, False, True),
    (3, 512, 128, False, True), (2, 200, 77, False, True), (1, 1024, 1024, True, False),
    (2, 384, 384, True, False), (2, 130, 130, True, False),
])
@pytest.mark.parametrize("workspace", [True, False])  # streaming dQ from stored dS^T vs the recomputing dQ kernel
def test_attention_fwd_bwd(cuda_dev, B, Lq, Lk, causal, masked, workspace, monkeypatch):
    from sct_gan_b200 import kernels as kn

    if not workspace:
        monkeypatch.setattr(kn, "_ATTN_WS_LIMIT", 0)

    H, dh = 8, 96
    d = H * dh
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + Lq + Lk)
    # packed projections as the model produces them: q from a [B*Lq, 3d] buffer when self-attention
    qkv = torch.randn(B * Lq, 3 * d, device="cuda", generator=g).to(BF16)
    q2 = qkv[:, :d]
    if Lq == Lk:
        k2, v2 = qkv[:, d:2 * d], qkv[:, 2 * d:]
    else:
        kv = torch.randn(B * Lk, 2 * d, device="cuda", generator=g).to(BF16)
        k2, v2 = kv[:, :d], kv[:, d:]
    kpm = None
    if masked:
        lens = torch.randint(max(1, Lk // 2), Lk + 1, (B,), device="cuda", generator=g)
        kpm = (torch.arange(Lk, device="cuda")[None, :] >= lens[:, None]).contiguous()
    o, lse2 = kn.attn_fwd(q2, k2, v2, B, H, Lq, Lk, kpm=kpm, causal=causal)
    _no_timeouts()

    def heads(t, L):
        return t.float().reshape(B, L, H, dh).permute(0, 2, 1, 3).contiguous().requires_grad_(True)

    qf, kf, vf = heads(q2, Lq), heads(k2, Lk), heads(v2, Lk)
    ref = _attn_ref(qf, kf, vf, kpm, causal, dh ** -0.5)
    o_h = o.float().reshape(B, Lq, H, dh).permute(0, 2, 1, 3)
    # P is rounded to bf16 before P@V and O is stored in bf16
    assert rel_l2(o_h, ref) < 1e-2

    d_o = torch.randn(B * Lq, d, device="cuda", generator=g).to(BF16)
    dqkv = torch.zeros(B * Lq, 3 * d, device="cuda", dtype=BF16)
    dq = dqkv[:, :d]
    if Lq == Lk:
        dk, dv = dqkv[:, d:2 * d], dqkv[:, 2 * d:]
    else:
        dkv = torch.zeros(B * Lk, 2 * d, device="cuda", dtype=BF16)
        dk, dv = dkv[:, :d], dkv[:, d:]
    kn.attn_bwd(q2, k2, v2, o, d_o, lse2, B, H, Lq, Lk, dq, dk, dv, kpm=kpm, causal=causal)
    _no_timeouts()
    ref.backward(d_o.float().reshape(B, Lq, H, dh).permute(0, 2, 1, 3))

    def unheads(t, L):
        return t.permute(0, 2, 1, 3).reshape(B * L, d)

    assert rel_l2(dv, unheads(vf.grad, Lk)) < 1.5e-2
    assert rel_l2(dk, unheads(kf.grad, Lk)) < 1.5e-2
    assert rel_l2(dq, unheads(qf.grad, Lq)) < 1.5e-2


@pytest.mark.parametrize("Lk,masked,t_max", [(1, False, 0), (77, True, 0), (300, True, 512), (1024, False, 1024),
                                              (513, True, 1024)])
def test_attention_decode_step(cuda_dev, Lk, masked, t_max):
    """Lq = 1 (the generation step) runs a SIMT kernel over the K/V cache: against fp32 softmax attention, with a
    key-padding mask, a cache whose batch stride exceeds Lk rows, and one fully masked row."""
    from sct_gan_b200 import kernels as kn

    B, H, dh = 5, 8, 96
    d = H * dh
    g = torch.Generator(device="cuda").manual_seed(Lk)
    q = torch.randn(B, d, device="cuda", generator=g).to(BF16)
    rows = t_max if t_max else Lk
    cache = torch.randn(B, rows, 2 * d, device="cuda", generator=g).to(BF16)
    kpm = None
    if masked:
        lens = torch.randint(1, Lk + 1, (B,), device="cuda", generator=g)
        lens[0] = 0  # every key masked: the output row is zero
        kpm = (torch.arange(Lk, device="cuda")[None, :] >= lens[:, None]).contiguous()
    flat = cache.view(B * rows, 2 * d)
    o, lse2 = kn.attn_fwd(q, flat[:, :d], flat[:, d:], B, H, 1, Lk, kpm=kpm, causal=False,
                          kv_batch_stride=rows * 2 * d if t_max else 0)
    kf = cache[:, :Lk, :d].float().reshape(B, Lk, H, dh).permute(0, 2, 1, 3)
    vf = cache[:, :Lk, d:].float().reshape(B, Lk, H, dh).permute(0, 2, 1, 3)
    qf = q.float().reshape(B, 1, H, dh).permute(0, 2, 1, 3)
    sc = torch.einsum("bhqd,bhkd->bhqk", qf, kf) * dh ** -0.5
    if kpm is not None:
        sc = sc.masked_fill(kpm[:, None, None, :], float("-inf"))
    ref = torch.nan_to_num(torch.softmax(sc, -1), nan=0.0) @ vf  # [B, H, 1, dh]
    got = o.float().reshape(B, 1, H, dh).permute(0, 2, 1, 3)
    assert (got - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    assert rel_l2(got, ref) < 5e-3
    if masked:
        assert torch.all(o[0] == 0) and torch.isinf(lse2[0]).all()
    ok = slice(1, None) if masked else slice(None)
    ref_lse = torch.logsumexp(sc, -1)[ok, :, 0] * 1.4426950408889634
    assert (lse2[ok, :, 0] - ref_lse).abs().max().item() < 1e-2


def test_attention_dropout_consistency(cuda_dev):
    """Dropout cannot match ATen's Philox stream; check the keep-rate/scale statistically and that the
    backward regenerates the same mask (finite-difference-free check: dV = P_drop^T dO is linear in dO)."""
    from sct_gan_b200 import kernels as kn

    B, H, dh, L = 2, 8, 96, 256
    d = H * dh
    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = (torch.randn(B * L, 3 * d, device="cuda", generator=g) * 0.1).to(BF16)
    q2, k2, v2 = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    v_ones = torch.ones_like(v2)
    qkv1 = torch.cat([q2, k2, v_ones], dim=1).contiguous()
    o, _ = kn.attn_fwd(qkv1[:, :d], qkv1[:, d:2 * d], qkv1[:, 2 * d:], B, H, L, L, p_drop=0.3, seed=11, offset=5)
    # with V = 1 every output element equals sum_k keep_k * p_k / 0.7: mean 1, never exactly 1
    assert abs(o.float().mean().item() - 1.0) < 0.02
    assert o.float().std().item() > 1e-3
    o2, _ = kn.attn_fwd(qkv1[:, :d], qkv1[:, d:2 * d], qkv1[:, 2 * d:], B, H, L, L, p_drop=0.3, seed=11, offset=5)
    assert torch.equal(o, o2)
    o3, _ = kn.attn_fwd(qkv1[:, :d], qkv1[:, d:2 * d], qkv1[:, 2 * d:], B, H, L, L, p_drop=0.3, seed=11, offset=6)
    assert not torch.equal(o, o3)


# ------------------------------------------------------------------------------------ row kernels
@pytest.mark.parametrize("B,S", [(2, 64), (3, 200), (8, 512)])
def test_embed_ln_pe(cuda_dev, B, S):
    from sct_gan_b200 import kernels as kn

    V, d = 50265, 768
    g = torch.Generator(device="cuda").manual_seed(S)
    ids = torch.randint(0, V, (B, S), device="cuda", generator=g)
    ids[0, :4] = torch.tensor([0, 1, 2, V - 1], device="cuda")
    ids[-1, -8:] = ids[0, 0:8]  # repeated ids: scatter-add collisions
    table = (torch.randn(V, d, device="cuda", generator=g) * 0.02).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(d, device="cuda", generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(d, device="cuda", generator=g)).requires_grad_(True)
    pos = torch.arange(1024, device="cuda", dtype=F32)[:, None]
    div = torch.exp(torch.arange(0, d, 2, device="cuda", dtype=F32) * (-math.log(10000.0) / d))
    pe = torch.zeros(1024, d, device="cuda")
    pe[:, 0::2], pe[:, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
    scale = math.sqrt(d)
    out_f32, out_bf16, stats = kn.embed_ln_pe_fwd(ids.view(-1), table.detach(), gamma.detach(), beta.detach(), pe, S, scale)
    x = torch.nn.functional.embedding(ids, table) * scale
    ref = torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-5) + pe[:S][None]
    # fp32 path: same arithmetic up to reduction order
    assert (out_f32.view(B, S, d) - ref).abs().max().item() < 2e-4
    assert torch.equal(out_bf16, out_f32.to(BF16))
    # gather index is integer work: with gamma=1, beta=0 un-normalising must give the exact table row
    gy = torch.randn(B * S, d, device="cuda", generator=g)
    dtable = torch.zeros_like(table)
    dgamma, dbeta = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    kn.embed_ln_pe_bwd(gy, ids.view(-1), table.detach(), gamma.detach(), stats, dtable, dgamma, dbeta, scale)
    ref.backward(gy.view(B, S, d))
    assert rel_l2(dtable, table.grad) < 1e-4
    assert rel_l2(dgamma, gamma.grad) < 1e-4
    assert rel_l2(dbeta, beta.grad) < 1e-4
    # rows never referenced stay exactly zero (bit-exact scatter targets)
    touched = torch.zeros(V, dtype=torch.bool, device="cuda")
    touched[ids.view(-1)] = True
    assert torch.equal((dtable.abs().sum(1) != 0) | ~touched, torch.ones_like(touched) & ((table.grad.abs().sum(1) != 0) | ~touched))
    assert dtable[~touched].abs().max().item() == 0.0
    # both outputs carried a gradient: the kernel sums the fp32 and the bf16 one while loading
    gy2 = torch.randn(B * S, d, device="cuda", generator=g).to(BF16)
    dt2 = torch.zeros_like(table)
    dg2, db2 = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    kn.embed_ln_pe_bwd(gy, ids.view(-1), table.detach(), gamma.detach(), stats, dt2, dg2, db2, scale, g2=gy2)
    dt3 = torch.zeros_like(table)
    dg3, db3 = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    kn.embed_ln_pe_bwd(gy + gy2.float(), ids.view(-1), table.detach(), gamma.detach(), stats, dt3, dg3, db3, scale)
    assert rel_l2(dt2, dt3) < 1e-6 and rel_l2(dg2, dg3) < 1e-5 and rel_l2(db2, db3) < 1e-5


def test_arena_buffers_are_adopted_by_autograd_without_a_copy(cuda_dev):
    """A parameter used at two sites of one step (the shared embedding table) accumulates into ONE arena buffer and
    autograd adopts that buffer as `.grad` (no clone): the gradient lives at the arena address."""
    from sct_gan_b200 import ops

    V, d, S = 1000, 768, 32
    table = (torch.randn(V, d, device="cuda") * 0.02).requires_grad_(True)
    gamma = torch.ones(d, device="cuda", requires_grad=True)
    beta = torch.zeros(d, device="cuda", requires_grad=True)
    pe = torch.zeros(S, d, device="cuda")
    ids = torch.randint(0, V, (2, S), device="cuda")

    def step():
        table.grad = gamma.grad = beta.grad = None
        ops.ARENA.begin(table.device)
        a, _ = ops.embed_ln_pe(ids, table, gamma, beta, pe, S, 1.0)
        b, _ = ops.embed_ln_pe(ids.flip(1), table, gamma, beta, pe, S, 1.0)
        (a.sum() + 2 * b.square().sum()).backward()
        ops.ARENA.end()

    step()  # sizing pass
    step()
    lo, hi = ops.ARENA.buf.data_ptr(), ops.ARENA.buf.data_ptr() + ops.ARENA.buf.numel() * 4
    assert lo <= table.grad.data_ptr() < hi and lo <= gamma.grad.data_ptr() < hi
    g_arena = table.grad.clone()
    ops.ARENA.buf = None
    ops.ARENA.need_last = 0
    table.grad = gamma.grad = beta.grad = None
    a, _ = ops.embed_ln_pe(ids, table, gamma, beta, pe, S, 1.0)
    b, _ = ops.embed_ln_pe(ids.flip(1), table, gamma, beta, pe, S, 1.0)
    (a.sum() + 2 * b.square().sum()).backward()
    assert rel_l2(g_arena, table.grad) < 1e-5


def test_embed_gather_bit_exact(cuda_dev):
    """With gamma=1, beta=0, pe=0 the output is LN(row*scale); un-normalising with the returned
    (mean, rstd) must reproduce the gathered table row for exactly the requested id."""
    from sct_gan_b200 import kernels as kn

    V, d = 1000, 768
    g = torch.Generator(device="cuda").manual_seed(3)
    table = torch.randn(V, d, device="cuda", generator=g)
    table[:, 0] = torch.arange(V, device="cuda", dtype=F32)  # id written into the row
    ids = torch.randint(0, V, (4 * 96,), device="cuda", generator=g)
    out, _, stats = kn.embed_ln_pe_fwd(ids, table, torch.ones(d, device="cuda"), torch.zeros(d, device="cuda"),
                                       torch.zeros(96, d, device="cuda"), 96, 1.0)
    rec = out / stats[:, 1:2] + stats[:, 0:1]
    assert torch.equal(rec[:, 0].round().long(), ids)


@pytest.mark.parametrize("d,n", [(768, 1000), (384, 1000), (1536, 1000), (768, 5003)])  # 5003 rows: staged kernels
def test_add_dropout_ln(cuda_dev, d, n):
    from sct_gan_b200 import kernels as kn

    g = torch.Generator(device="cuda").manual_seed(d)
    x = torch.randn(n, d, device="cuda", generator=g).requires_grad_(True)
    br = torch.randn(n, d, device="cuda", generator=g).to(BF16)
    brf = br.float().requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(d, device="cuda", generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(d, device="cuda", generator=g)).requires_grad_(True)
    x_out, y_ln, y_cast, stats = kn.add_dropout_ln_fwd(x.detach(), br, 0.1, gamma.detach(), beta.detach(), want_cast=True)
    xr = x + 0.1 * brf
    yr = torch.nn.functional.layer_norm(xr, (d,), gamma, beta, 1e-5)
    assert (x_out - xr).abs().max().item() < 1e-5
    assert rel_l2(y_ln, yr) < 4e-3  # bf16 output
    assert torch.equal(y_cast, x_out.to(BF16))
    g_x = torch.randn(n, d, device="cuda", generator=g)
    g_y = torch.randn(n, d, device="cuda", generator=g).to(BF16)
    dgamma, dbeta = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    gx, gb = kn.add_dropout_ln_bwd(g_x, g_y, None, x_out, stats, gamma.detach(), 0.1, dgamma, dbeta)
    (xr * g_x).sum().backward(retain_graph=True)
    (yr * g_y.float()).sum().backward()
    assert rel_l2(gx, x.grad) < 1e-4
    assert rel_l2(gb, brf.grad) < 4e-3
    assert rel_l2(dgamma, gamma.grad) < 1e-3
    assert rel_l2(dbeta, beta.grad) < 1e-3
    if n >= 4096 and d == 768:
        # the software-pipelined backward (>= 4096 rows) against the plain kernel (fewer rows) on the same leading rows,
        # with all three incoming gradients and dropout: identical per-row arithmetic, so identical bits
        g_c = torch.randn(n, d, device="cuda", generator=g).to(BF16)
        m = 4000
        dg1, db1 = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
        gx1, gb1 = kn.add_dropout_ln_bwd(g_x, g_y, g_c, x_out, stats, gamma.detach(), 0.1, dg1, db1, p_drop=0.3, seed=3, offset=4)
        dg2, db2 = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
        gx2, gb2 = kn.add_dropout_ln_bwd(g_x[:m], g_y[:m], g_c[:m], x_out[:m], stats[:m], gamma.detach(), 0.1, dg2, db2,
                                         p_drop=0.3, seed=3, offset=4)
        assert torch.equal(gx1[:m], gx2) and torch.equal(gb1[:m], gb2)
        dg3, db3 = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
        kn.add_dropout_ln_bwd(g_x[m:], g_y[m:], g_c[m:], x_out[m:], stats[m:], gamma.detach(), 0.1, dg3, db3, p_drop=0.3,
                              seed=3, offset=4)
        assert rel_l2(dg1, dg2 + dg3) < 1e-5 and rel_l2(db1, db2 + db3) < 1e-5


@pytest.mark.parametrize("n", [2048, 4500])  # plain and bulk-copy-staged kernels generate the same masks
def test_dropout_mask_replay(cuda_dev, n):
    from sct_gan_b200 import kernels as kn

    d = 768
    x = torch.zeros(n, d, device="cuda")
    br = torch.ones(n, d, device="cuda", dtype=BF16)
    x_out, _, _, _ = kn.add_dropout_ln_fwd(x, br, 1.0, None, None, want_ln=False, p_drop=0.3, seed=5, offset=9)
    keep = x_out != 0
    assert abs(keep.float().mean().item() - 0.7) < 5e-3
    assert torch.allclose(x_out[keep], torch.full_like(x_out[keep], 1 / 0.7), rtol=1e-6)
    gx, gb = kn.add_dropout_ln_bwd(torch.ones(n, d, device="cuda"), None, None, None, None, None, 1.0, None, None,
                                   p_drop=0.3, seed=5, offset=9)
    assert torch.equal(gb.float() != 0, keep)  # backward regenerates the forward's mask


@pytest.mark.parametrize("d", [768, 384, 1536])
def test_ln_act(cuda_dev, d):
    from sct_gan_b200 import kernels as kn

    n = 777
    g = torch.Generator(device="cuda").manual_seed(d + 1)
    z = torch.randn(n, d, device="cuda", generator=g).to(BF16)
    zf = z.float().requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(d, device="cuda", generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(d, device="cuda", generator=g)).requires_grad_(True)
    h, stats = kn.ln_act_fwd(z, gamma.detach(), beta.detach())
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(zf, (d,), gamma, beta, 1e-5))
    assert rel_l2(h, ref) < 4e-3
    gh = torch.randn(n, d, device="cuda", generator=g).to(BF16)
    dgamma, dbeta = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    gz = kn.ln_act_bwd(gh, z, stats, gamma.detach(), beta.detach(), dgamma, dbeta)
    ref.backward(gh.float())
    assert rel_l2(gz, zf.grad) < 5e-3
    assert rel_l2(dgamma, gamma.grad) < 2e-3
    assert rel_l2(dbeta, beta.grad) < 2e-3


def test_gelu_dropout(cuda_dev):
    from sct_gan_b200 import kernels as kn

    g = torch.Generator(device="cuda").manual_seed(1)
    z = (2 * torch.randn(512, 2048, device="cuda", generator=g)).to(BF16)
    zf = z.float().requires_grad_(True)
    h = kn.gelu_dropout_fwd(z)
    ref = torch.nn.functional.gelu(zf)
    assert rel_l2(h, ref) < 4e-3
    gh = torch.randn(512, 2048, device="cuda", generator=g).to(BF16)
    gz = kn.gelu_dropout_bwd(gh, z)
    ref.backward(gh.float())
    assert rel_l2(gz, zf.grad) < 5e-3


def test_colsum_cast_seqmean(cuda_dev):
    from sct_gan_b200 import kernels as kn

    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(3000, 776, device="cuda", generator=g).to(BF16)
    out = torch.ones(776, device="cuda")
    kn.colsum_bf16(x, out, 0.5)
    assert rel_l2(out, 1 + 0.5 * x.float().sum(0)) < 1e-4
    src = torch.randn(100, 768, device="cuda", generator=g)
    dst = torch.zeros(100, 1536, device="cuda", dtype=BF16)
    kn.cast_scale(src, dst, col_off=768, scale=0.1)
    assert torch.equal(dst[:, 768:], (src * 0.1).to(BF16)) and dst[:, :768].abs().max().item() == 0
    B, S, d = 4, 300, 768
    xf = torch.randn(B * S, d, device="cuda", generator=g)
    yb = torch.randn(B * S, d, device="cuda", generator=g).to(BF16)
    m = kn.seq_mean_fwd(xf, yb, B, S, d)
    assert rel_l2(m, (xf + yb.float()).view(B, S, d).mean(1)) < 1e-5
    gx, gy = kn.seq_mean_bwd(m, B, S, d, want_f32=True, want_bf16=True)
    assert rel_l2(gx.view(B, S, d), (m / S)[:, None, :].expand(B, S, d)) < 1e-6
    assert torch.equal(gy, gx.to(BF16))


# ------------------------------------------------------------------------------------ loss kernels
@pytest.mark.parametrize("rows,V", [(64, 50265), (257, 1000), (16, 8)])
def test_ce_rows(cuda_dev, rows, V):
    from sct_gan_b200 import kernels as kn

    g = torch.Generator(device="cuda").manual_seed(V)
    ld = (V + 7) // 8 * 8
    logits = (3 * torch.randn(rows, ld, device="cuda", generator=g)).to(BF16)
    tgt = torch.randint(0, V, (rows,), device="cuda", generator=g)
    tgt[1] = -1  # excluded row
    lf = logits[:, :V].float().requires_grad_(True)
    work = logits.clone()
    row_loss, row_lse = kn.ce_rows(work[:, :V], tgt, V, grad_scale=1.0 / rows, write_grad=True)
    valid = tgt >= 0
    ref_rows = torch.nn.functional.cross_entropy(lf[valid], tgt[valid], reduction="none")
    assert (row_loss[valid] - ref_rows).abs().max().item() < 2e-5 * max(1.0, ref_rows.abs().max().item()) + 1e-5
    assert row_loss[1].item() == 0.0
    (ref_rows.sum() / rows).backward()
    gref = lf.grad
    assert rel_l2(work[:, :V][valid], gref[valid]) < 5e-3  # bf16 gradient storage
    assert work[1, :V].abs().max().item() == 0.0


@pytest.mark.parametrize("K,N", [(768, 1536), (1536, 768), (768, 384), (384, 1)])
def test_small_linear(cuda_dev, K, N):
    from sct_gan_b200 import kernels as kn

    M = 32
    g = torch.Generator(device="cuda").manual_seed(K + N)
    x = torch.randn(M, K, device="cuda", generator=g).requires_grad_(True)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).requires_grad_(True)
    b = torch.randn(N, device="cuda", generator=g).requires_grad_(True)
    y = kn.small_linear_fwd(x.detach(), w.detach(), b.detach())
    ref = torch.nn.functional.linear(x, w, b)
    assert rel_l2(y, ref) < 1e-5
    dy = torch.randn(M, N, device="cuda", generator=g)
    dx, dw, db = kn.small_linear_bwd(dy, x.detach(), w.detach())
    ref.backward(dy)
    assert rel_l2(dx, x.grad) < 1e-5 and rel_l2(dw, w.grad) < 1e-5 and rel_l2(db, b.grad) < 1e-5


@pytest.mark.parametrize("shift", [-3.0, 0.0, 3.0])
def test_gan_loss(cuda_dev, shift):
    """train.py:1201-1234 — thresholds 0.3 / 0.8 exercised by shifting the logits."""
    from sct_gan_b200 import kernels as kn

    g = torch.Generator(device="cuda").manual_seed(11)
    z = (torch.randn(32, device="cuda", generator=g) + shift).requires_grad_(True)
    out4 = kn.gan_loss_fwd(z.detach())
    bce = torch.nn.BCEWithLogitsLoss()
    zz = z.view(-1, 1)
    d = bce(zz, torch.ones_like(zz))
    c = torch.sigmoid(zz).mean().item()
    adv = bce(zz, torch.zeros_like(zz)) if c < 0.3 else torch.zeros((), device="cuda")
    if c > 0.8:
        d = d + 1.0 * torch.mean(torch.sigmoid(zz) ** 2) + 2.0 * torch.mean(torch.sigmoid(zz) ** 4)
    assert abs(out4[0].item() - d.item()) < 1e-5 and abs(out4[1].item() - adv.item()) < 1e-5
    assert abs(out4[2].item() - c) < 1e-6
    (0.05 * d + 0.02 * adv).backward()
    dz = kn.gan_loss_bwd(z.detach(), out4[2:3], torch.tensor([0.05], device="cuda"), torch.tensor([0.02], device="cuda"))
    assert rel_l2(dz, z.grad) < 1e-4


@pytest.mark.parametrize("causal", [False, True])
def test_attention_dropout_mask_fwd_bwd_identical(cuda_dev, causal):
    """The three attention kernels regenerate the same dropout mask: extract it from the forward (V = identity
    blocks make O = dropout(P)), then check dQ/dK/dV against PyTorch autograd with that explicit mask."""
    from sct_gan_b200 import kernels as kn

    B, H, dh, L, p_drop = 2, 8, 96, 192, 0.3
    d = H * dh
    g = torch.Generator(device="cuda").manual_seed(17)
    q2 = (torch.randn(B * L, d, device="cuda", generator=g) * 0.5).to(BF16)
    kv = (torch.randn(B * L, 2 * d, device="cuda", generator=g) * 0.5).to(BF16)  # packed k|v as the model has them
    k2 = kv[:, :d]
    eye = torch.eye(dh, device="cuda", dtype=BF16)
    keep = torch.zeros(B, H, L, L, device="cuda", dtype=torch.bool)
    for blk in range(L // dh):  # O = dropout(P)[:, :, :, blk*96:(blk+1)*96] when V is the identity on that block
        v = torch.zeros(B, L, H, dh, device="cuda", dtype=BF16)
        v[:, blk * dh:(blk + 1) * dh] = eye[None, :, None, :].expand(B, dh, H, dh)
        kv[:, d:] = v.reshape(B * L, d)
        o, _ = kn.attn_fwd(q2, k2, kv[:, d:], B, H, L, L, causal=causal, p_drop=p_drop, seed=3, offset=8)
        keep[:, :, :, blk * dh:(blk + 1) * dh] = o.reshape(B, L, H, dh).permute(0, 2, 1, 3) != 0
    if causal:
        tri = torch.ones(L, L, device="cuda", dtype=torch.bool).tril()
        assert not keep[:, :, ~tri].any()
        rate = keep[:, :, tri].float().mean().item()
    else:
        rate = keep.float().mean().item()
    assert abs(rate - 0.7) < 5e-3, rate
    kv[:, d:] = torch.randn(B * L, d, device="cuda", generator=g).to(BF16)
    v2 = kv[:, d:]
    o, lse2 = kn.attn_fwd(q2, k2, v2, B, H, L, L, causal=causal, p_drop=p_drop, seed=3, offset=8)

    def heads(t):
        return t.float().reshape(B, L, H, dh).permute(0, 2, 1, 3).contiguous().requires_grad_(True)

    qf, kf, vf = heads(q2), heads(k2), heads(v2)
    sc = torch.einsum("bhqd,bhkd->bhqk", qf, kf) * dh ** -0.5
    if causal:
        sc = sc.masked_fill(~tri, float("-inf"))
    pr = torch.softmax(sc, -1) * keep / (1 - p_drop)
    ref = pr @ vf
    assert rel_l2(o.float().reshape(B, L, H, dh).permute(0, 2, 1, 3), ref) < 1e-2
    d_o = torch.randn(B * L, d, device="cuda", generator=g).to(BF16)
    dq = torch.zeros(B * L, d, device="cuda", dtype=BF16)
    dkv = torch.zeros(B * L, 2 * d, device="cuda", dtype=BF16)
    dk, dv = dkv[:, :d], dkv[:, d:]
    kn.attn_bwd(q2, k2, v2, o, d_o, lse2, B, H, L, L, dq, dk, dv, causal=causal, p_drop=p_drop, seed=3, offset=8)
    ref.backward(d_o.float().reshape(B, L, H, dh).permute(0, 2, 1, 3))

    def unheads(t):
        return t.permute(0, 2, 1, 3).reshape(B * L, d)

    assert rel_l2(dv, unheads(vf.grad)) < 1.5e-2
    assert rel_l2(dk, unheads(kf.grad)) < 1.5e-2
    assert rel_l2(dq, unheads(qf.grad)) < 1.5e-2


# ------------------------------------------------------------------------------------------------
# fused GEMM epilogues of the feed-forward block
@pytest.mark.parametrize("M,N,K", [(1024, 2048, 768), (300, 512, 768), (4096 + 77, 2048, 768)])
@pytest.mark.parametrize("p_drop", [0.0, 0.3])
def test_gemm_fused_gelu_epilogues(cuda_dev, M, N, K, p_drop):
    """sct_gemm_bf16_nt_gelu / sct_gemm_bf16_nn_mul against the unfused fp32 PyTorch chain linear -> F.gelu (exact
    erf, as the reference's activation='gelu') -> dropout and its backward.  The dropout mask cannot be matched to
    ATen's Philox stream: it is read off the kernel's own output (h == 0 where gelu(z) != 0), checked statistically,
    for independence across rows and for h / g consistency.  Tolerance: bf16 outputs, tanh-fit GELU (2.7e-4 abs)."""
    from sct_gan_b200 import kernels as kn

    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w1 = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b1 = torch.randn(N, device="cuda", generator=g) * 0.1
    h, gd = kn.gemm_nt_gelu(x, w1, b1, p_drop, seed=7, offset=3)
    zf = (x.float() @ w1.float().t() + b1).requires_grad_(True)
    act = torch.nn.functional.gelu(zf)  # exact erf form
    act.sum().backward()
    g_ref, d_ref = act.detach(), zf.grad
    keep = 1.0 - (77.0 / 256.0 if p_drop > 0 else 0.0)  # p is resolved to 2^-8; the scale uses the realised value
    live = g_ref.abs() > 2e-2
    if p_drop == 0.0:
        mask = torch.ones_like(g_ref, dtype=torch.bool)
    else:
        mask = h.float().abs() > 0
        frac = mask[live].float().mean().item()
        assert abs(frac - keep) < 5e-3, frac
        per_row = (mask & live).float().sum(dim=1) / live.float().sum(dim=1).clamp(min=1)
        assert per_row.min().item() > keep - 0.2 and per_row.max().item() < keep + 0.2  # rows draw independently
        h2, _ = kn.gemm_nt_gelu(x, w1, b1, p_drop, seed=7, offset=3)
        assert torch.equal(h, h2)
        h3, _ = kn.gemm_nt_gelu(x, w1, b1, p_drop, seed=7, offset=4)
        assert not torch.equal(h, h3)
        # h and g carry the same mask
        assert torch.equal(mask & live & (d_ref.abs() > 2e-2), (gd.float().abs() > 0) & live & (d_ref.abs() > 2e-2))
    mf = mask.float() / keep
    assert rel_l2(h.float() * live, g_ref * mf * live) < 1e-2
    assert (h.float() - g_ref * mf)[mask].abs().max().item() < 2e-2 * max(1.0, g_ref.abs().max().item() / keep)
    assert rel_l2(gd.float() * mask, d_ref * mf) < 1e-2
    assert (gd.float() - d_ref * mf)[mask].abs().max().item() < 2e-2 / keep
    # backward GEMM: dz = (dy @ w2) * g
    w2 = (torch.randn(K, N, device="cuda", generator=g) * 0.05).bfloat16()  # linear2.weight [K_out = K, N]
    dy = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    dz = kn.gemm_nn_mul(dy, w2, gd)
    dz_ref = (dy.float() @ w2.float()) * gd.float()
    assert rel_l2(dz, dz_ref) < 1e-2
    _no_timeouts()


def test_fused_ffn_autograd_matches_unfused(cuda_dev):
    """ops.fused_ffn (4 GEMMs, fused epilogues) against the unfused kernels path at p = 0: outputs and all gradients."""
    import torch.nn as nn

    from sct_gan_b200 import ops

    torch.manual_seed(0)
    l1, l2 = nn.Linear(768, 2048).cuda(), nn.Linear(2048, 768).cuda()
    y = (torch.randn(1000, 768, device="cuda") * 0.7).bfloat16()
    w1b, w2b = l1.weight.detach().bfloat16(), l2.weight.detach().bfloat16()
    outs = []
    for fused in (True, False):
        for p in (*l1.parameters(), *l2.parameters()):
            p.grad = None
        yy = y.clone().requires_grad_(True)
        if fused:
            o = ops.fused_ffn(yy, l1, l2, w1b, w2b, 0.0)
        else:
            o = ops.linear(ops.gelu_dropout(ops.linear(yy, l1.weight, l1.bias, w1b), 0.0), l2.weight, l2.bias, w2b)
        (o.float() * torch.linspace(-1, 1, 768, device="cuda")).sum().backward()
        outs.append([o.float(), yy.grad.float(), l1.weight.grad.clone(), l1.bias.grad.clone(), l2.weight.grad.clone(),
                     l2.bias.grad.clone()])
    for a, b in zip(*outs):
        assert ((a - b).norm() / b.norm()).item() < 1e-2


@pytest.mark.parametrize("causal", [False, True])
def test_attention_fwd_running_max_rescale(cuda_dev, causal):
    """Scores that GROW along the key axis (keys scaled by a ramp): every later tile raises the row maximum by far
    more than the lazy-rescale threshold (2^8), so the forward's reference-maximum update and the rescale of the O
    accumulator in TMEM run on every tile — including rows of a warp that do not move while others do.  Rows are
    given different growth rates so the rescale is divergent inside warps."""
    from sct_gan_b200 import kernels as kn

    B, H, dh, L = 2, 8, 96, 768
    d = H * dh
    g = torch.Generator(device="cuda").manual_seed(77)
    q = torch.randn(B * L, d, device="cuda", generator=g)
    k = torch.randn(B * L, d, device="cuda", generator=g)
    v = torch.randn(B * L, d, device="cuda", generator=g)
    # align k with a fixed direction and ramp its length along the sequence; query rows project on that direction with
    # row-dependent strength (some negative: their scores shrink instead, so their reference never moves)
    u = torch.nn.functional.normalize(torch.randn(dh, device="cuda", generator=g), dim=0)
    ramp = torch.linspace(0.0, 300.0, L, device="cuda").repeat(B)[:, None]
    k = (k.view(B * L, H, dh) * 0.3 + ramp[:, None, :] * u).reshape(B * L, d)
    gain = torch.linspace(-1.0, 3.0, L, device="cuda").repeat(B)[:, None]
    q = (q.view(B * L, H, dh) * 0.3 + gain[:, None, :] * u).reshape(B * L, d)
    qb, kb, vb = q.to(BF16), k.to(BF16), v.to(BF16)
    o, lse2 = kn.attn_fwd(qb, kb, vb, B, H, L, L, causal=causal)
    _no_timeouts()

    def heads(t):
        return t.float().reshape(B, L, H, dh).permute(0, 2, 1, 3)

    ref = _attn_ref(heads(qb), heads(kb), heads(vb), None, causal, dh ** -0.5)
    assert rel_l2(o.float().reshape(B, L, H, dh).permute(0, 2, 1, 3), ref) < 1e-2
    s = torch.einsum("bhqd,bhkd->bhqk", heads(qb), heads(kb)) * dh ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(L, L, device="cuda", dtype=torch.bool).triu(1), float("-inf"))
    lse_ref = torch.logsumexp(s, dim=-1) * math.log2(math.e)
    assert (lse2 - lse_ref).abs().max().item() < 5e-2 * max(1.0, lse_ref.abs().max().item())
    assert float((s.max(dim=-1).values * math.log2(math.e)).max()) > 60  # the growth really exceeded the threshold


# ------------------------------------------------------------------------------------ sampling tail (§8 f4)
def _kept_set(row_f32, tweak):
    """Token ids the reference rule (model.py:892-918) can emit for one row, ties to the lowest index."""
    z = row_f32.double().clone()
    z /= 0.7
    if tweak:
        z[59] *= 2.0
    srt = torch.sort(z, descending=True, stable=True)
    k = min(50, z.numel())
    topv, topi = srt.values[:k], srt.indices[:k]
    pr = torch.softmax(topv, dim=-1)
    remove = torch.cumsum(pr, dim=-1) > 0.95
    remove[1:] = remove[:-1].clone()
    remove[0] = False
    return set(topi[~remove].tolist()), int(topi[0])


@pytest.mark.parametrize("V", [50265, 40, 4096])
def test_sample_rows_kernel(cuda_dev, V):
    from sct_gan_b200 import kernels as kn

    B = 24
    g = torch.Generator().manual_seed(V)
    ld = (V + 7) // 8 * 8
    rows = (torch.randn(B, V, generator=g) * 2.0).bfloat16()
    if V >= 64:
        rows[3] = 0.5                  # a constant row: 50 survivors are ids 0..49 (ties -> lowest index), argmax = 0
    rows[4, 7], rows[4, V - 1] = 30.0, 30.0   # the maximum appears twice: argmax = 7
    if V > 59:
        rows[5] = 0.0
        rows[5, 59], rows[5, 10] = 3.0, 4.0   # after the x2 tweak id 59 (6.0) beats id 10 (4.0)
    buf = torch.full((B, ld), float("inf"), dtype=BF16, device="cuda")  # pitch columns hold garbage larger than any logit
    buf[:, :V] = rows.cuda()
    logits = buf[:, :V]
    prev = torch.full((B,), 7, dtype=torch.long, device="cuda")
    prev[5], prev[6] = 2001, 2002
    # greedy = lowest index among the maxima of the (tweaked) row
    got = kn.sample_rows(logits, V, prev, greedy=True).view(-1).cpu()
    for b in range(B):
        tweak = V > 59 and int(prev[b]) in (2000, 2001, 2002)
        assert int(got[b]) == _kept_set(rows[b].float(), tweak)[1], b
    assert torch.equal(kn.sample_rows(logits, V, None, greedy=True).view(-1).cpu()[:5], got[:5])
    # sampling: every draw lies inside the reference's nucleus; the counter changes the draw, the same counter repeats it
    ep = torch.zeros(1, dtype=torch.long, device="cuda")
    seen = [set() for _ in range(B)]
    first = None
    for it in range(60):
        ep.fill_(it)
        nxt = kn.sample_rows(logits, V, prev, 0.7, 50, 0.95, False, seed=11, offset=3, epoch=ep).view(-1).cpu()
        first = nxt if first is None else first
        for b in range(B):
            seen[b].add(int(nxt[b]))
    for b in range(B):
        tweak = V > 59 and int(prev[b]) in (2000, 2001, 2002)
        kept, _ = _kept_set(rows[b].float(), tweak)
        assert seen[b] <= kept, (b, sorted(seen[b] - kept))
    if V >= 64:
        assert len(seen[3]) > 10 and seen[3] <= set(range(48))  # the constant row samples among its 48 kept ids
    ep.fill_(0)
    again = kn.sample_rows(logits, V, prev, 0.7, 50, 0.95, False, seed=11, offset=3, epoch=ep).view(-1).cpu()
    assert torch.equal(again, first)
    other = kn.sample_rows(logits, V, prev, 0.7, 50, 0.95, False, seed=12, offset=3, epoch=ep).view(-1).cpu()
    assert not torch.equal(other, first)
