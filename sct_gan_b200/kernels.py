"""Raw (non-autograd) Python entry points of the C ABI: torch tensors in, torch tensors out.

Each function validates dtype / contiguity / device, then hands raw device pointers and the current
CUDA stream to libsct_b200.so (include/sct_b200.h).  PyTorch owns all memory; there is no fallback.
The autograd layer (ops.py) and the tests call these.
"""
from __future__ import annotations

import os

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t, align=16):
    if t is None:
        return None
    assert t.is_cuda, "sct_gan_b200 kernels need CUDA tensors (no CPU fallback)"
    assert t.data_ptr() % align == 0, f"tensor storage must be {align}-byte aligned"
    return t.data_ptr()


def _eptr(t):
    """Per-call dropout epoch: a device int64 tensor the kernel reads at run time (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.int64
    return t.data_ptr()


def _sptr(t):
    """Pointer to a device scalar / small fp32 vector (4-byte alignment is enough)."""
    return _ptr(t, 4)


def _chk(t, dtype, name):
    assert t.dtype == dtype, f"{name}: expected {dtype}, got {t.dtype}"
    assert t.is_contiguous(), f"{name}: must be contiguous"
    return t


def _rows2d(t, dtype, name):
    """[rows, cols] view with unit inner stride; returns (tensor, ld)."""
    assert t.dtype == dtype, f"{name}: expected {dtype}, got {t.dtype}"
    assert t.dim() == 2 and t.stride(1) == 1, f"{name}: need a 2-D row-major view"
    return t, t.stride(0)


# ---------------------------------------------------------------------------------------------- K1
def embed_ln_pe_fwd(ids, table, gamma, beta, pe, seq_len, scale, p_drop=0.0, seed=0, offset=0, epoch=None,
                    want_f32=True, want_bf16=True):
    ids = _chk(ids, torch.int64, "ids")
    n_tok = ids.numel()
    vocab, d = table.shape
    _chk(table, F32, "table"), _chk(gamma, F32, "gamma"), _chk(beta, F32, "beta"), _chk(pe, F32, "pe")
    assert pe.numel() >= seq_len * d
    dev = ids.device
    out_f32 = torch.empty((n_tok, d), dtype=F32, device=dev) if want_f32 else None
    out_bf16 = torch.empty((n_tok, d), dtype=BF16, device=dev) if want_bf16 else None
    stats = torch.empty((n_tok, 2), dtype=F32, device=dev)
    # algorithmic bytes per token: id + fp32 table row + outputs + (mean, rstd)
    _lib.Stats.annotate(float(n_tok) * (8 + 4 * d + (4 * d if want_f32 else 0) + (2 * d if want_bf16 else 0) + 8))
    _lib.call("sct_embed_ln_pe_fwd", _ptr(ids), _ptr(table), _ptr(gamma), _ptr(beta), _ptr(pe),
              _ptr(out_f32), _ptr(out_bf16), _ptr(stats), n_tok, seq_len, vocab, d, float(scale),
              float(p_drop), seed, offset, _eptr(epoch), _stream())
    return out_f32, out_bf16, stats


def embed_ln_pe_bwd(g, ids, table, gamma, stats, dtable, dgamma, dbeta, scale, p_drop=0.0, seed=0,
                    offset=0, epoch=None, g2=None):
    """dtable / dgamma / dbeta (fp32) are accumulated into.  `g` and the optional `g2` are the incoming gradients of
    the fp32 and the bf16 output (one of each dtype at most); the kernel sums them while loading."""
    vocab, d = table.shape
    n_tok = ids.numel()
    gs = [t for t in (g, g2) if t is not None]
    g_f32 = next((t for t in gs if t.dtype == F32), None)
    g_bf16 = next((t for t in gs if t.dtype == BF16), None)
    assert len(gs) == (g_f32 is not None) + (g_bf16 is not None), "at most one gradient per dtype"
    for t in gs:
        assert t.is_contiguous() and t.numel() == n_tok * d
    # dY + re-gathered fp32 row + fp32 read-modify-write scatter into the table gradient + stats
    _lib.Stats.annotate(float(n_tok) * (sum(t.element_size() for t in gs) * d + 4 * d + 8 * d + 8 + 8))
    _lib.call("sct_embed_ln_pe_bwd", _ptr(g_f32), _ptr(g_bf16), _ptr(ids), _ptr(table), _ptr(gamma),
              _ptr(stats), _ptr(dtable), _ptr(dgamma), _ptr(dbeta), n_tok, vocab, d, float(scale),
              float(p_drop), seed, offset, _eptr(epoch), _stream())


# --------------------------------------------------------------------------------------------- K4a
def add_dropout_ln_fwd(x, branch, alpha, gamma, beta, want_x=True, want_ln=True, want_cast=False,
                       p_drop=0.0, seed=0, offset=0, epoch=None):
    ref = x if x is not None else branch
    n_rows, d = ref.shape
    dev = ref.device
    if x is not None:
        _chk(x, F32, "x")
    if branch is not None:
        _chk(branch, BF16, "branch")
    x_out = torch.empty((n_rows, d), dtype=F32, device=dev) if want_x else None
    y_ln = torch.empty((n_rows, d), dtype=BF16, device=dev) if want_ln else None
    y_cast = torch.empty((n_rows, d), dtype=BF16, device=dev) if want_cast else None
    stats = torch.empty((n_rows, 2), dtype=F32, device=dev) if want_ln else None
    _lib.Stats.annotate(float(n_rows * d * ((4 if x is not None else 0) + (2 if branch is not None else 0)
                                            + (4 if want_x else 0) + (2 if want_ln else 0) + (2 if want_cast else 0))))
    _lib.call("sct_add_dropout_ln_fwd", _ptr(x), _ptr(branch), float(alpha), _ptr(gamma), _ptr(beta),
              _ptr(x_out), _ptr(y_ln), _ptr(y_cast), _ptr(stats), n_rows, d, float(p_drop), seed,
              offset, _eptr(epoch), _stream())
    return x_out, y_ln, y_cast, stats


def add_dropout_ln_bwd(g_xout, g_yln, g_ycast, xprime, stats, gamma, alpha, dgamma, dbeta,
                       want_gx=True, want_gbranch=True, p_drop=0.0, seed=0, offset=0, epoch=None):
    ref = next(t for t in (g_xout, g_yln, g_ycast) if t is not None)
    n_rows, d = ref.shape
    dev = ref.device
    g_x = torch.empty((n_rows, d), dtype=F32, device=dev) if want_gx else None
    g_branch = torch.empty((n_rows, d), dtype=BF16, device=dev) if want_gbranch else None
    _lib.Stats.annotate(float(n_rows * d) * ((4 if g_xout is not None else 0) + (2 if g_yln is not None else 0)
                                             + (2 if g_ycast is not None else 0) + (4 if xprime is not None else 0)
                                             + (4 if want_gx else 0) + (2 if want_gbranch else 0)))
    _lib.call("sct_add_dropout_ln_bwd", _ptr(g_xout), _ptr(g_yln), _ptr(g_ycast), _ptr(xprime),
              _ptr(stats), _ptr(gamma), float(alpha), _ptr(g_x), _ptr(g_branch), _ptr(dgamma),
              _ptr(dbeta), n_rows, d, float(p_drop), seed, offset, _eptr(epoch), _stream())
    return g_x, g_branch


def ln_act_fwd(z, gamma, beta, p_drop=0.0, seed=0, offset=0, epoch=None):
    _chk(z, BF16, "z")
    n_rows, d = z.shape
    h = torch.empty_like(z)
    stats = torch.empty((n_rows, 2), dtype=F32, device=z.device)
    _lib.Stats.annotate(float(n_rows * d) * 4)
    _lib.call("sct_ln_act_fwd", _ptr(z), _ptr(gamma), _ptr(beta), _ptr(h), _ptr(stats), n_rows, d,
              float(p_drop), seed, offset, _eptr(epoch), _stream())
    return h, stats


def ln_act_bwd(g_h, z, stats, gamma, beta, dgamma, dbeta, p_drop=0.0, seed=0, offset=0, epoch=None):
    _chk(g_h, BF16, "g_h")
    n_rows, d = z.shape
    g_z = torch.empty_like(z)
    _lib.Stats.annotate(float(n_rows * d) * 6)
    _lib.call("sct_ln_act_bwd", _ptr(g_h), _ptr(z), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(g_z),
              _ptr(dgamma), _ptr(dbeta), n_rows, d, float(p_drop), seed, offset, _eptr(epoch), _stream())
    return g_z


def gelu_dropout_fwd(z, p_drop=0.0, seed=0, offset=0, epoch=None):
    _chk(z, BF16, "z")
    h = torch.empty_like(z)
    _lib.Stats.annotate(float(z.numel()) * 4)
    _lib.call("sct_gelu_dropout_fwd", _ptr(z), _ptr(h), z.numel(), float(p_drop), seed, offset, _eptr(epoch), _stream())
    return h


def gelu_dropout_bwd(g_h, z, p_drop=0.0, seed=0, offset=0, epoch=None):
    _chk(g_h, BF16, "g_h"), _chk(z, BF16, "z")
    g_z = torch.empty_like(z)
    _lib.Stats.annotate(float(z.numel()) * 6)
    _lib.call("sct_gelu_dropout_bwd", _ptr(g_h), _ptr(z), _ptr(g_z), z.numel(), float(p_drop), seed,
              offset, _eptr(epoch), _stream())
    return g_z


def colsum_bf16(x, out, scale=1.0):
    """out[n] += scale * sum_m x[m, n]"""
    x, ld = _rows2d(x, BF16, "x")
    M, N = x.shape
    _chk(out, F32, "out")
    _lib.Stats.annotate(float(M * N) * 2)
    _lib.call("sct_colsum_bf16", _ptr(x), ld, _ptr(out), M, N, float(scale), _stream())


def cast_scale(src, dst, col_off=0, scale=1.0):
    """dst[:, col_off:col_off+cols] = bf16(scale * src)"""
    rows, cols = src.shape
    assert src.dim() == 2 and src.stride(1) == 1
    dst, ld = _rows2d(dst, BF16, "dst")
    f = src if src.dtype == F32 else None
    b = src if src.dtype == BF16 else None
    _lib.Stats.annotate(float(rows * cols) * (src.element_size() + 2))
    _lib.call("sct_cast_scale", _ptr(f, 8), _ptr(b, 8), src.stride(0), _ptr(dst, 8), rows, cols, ld, col_off,
              float(scale), _stream())


def seq_mean_fwd(x, y, B, S, d):
    ref = x if x is not None else y
    out = torch.empty((B, d), dtype=F32, device=ref.device)
    _lib.Stats.annotate(float(B * S * d) * ((4 if x is not None else 0) + (2 if y is not None else 0)))
    _lib.call("sct_seq_mean_fwd", _ptr(x), _ptr(y), _ptr(out), B, S, d, _stream())
    return out


def seq_mean_bwd(g, B, S, d, want_f32=False, want_bf16=True):
    _chk(g, F32, "g")
    gx = torch.empty((B * S, d), dtype=F32, device=g.device) if want_f32 else None
    gy = torch.empty((B * S, d), dtype=BF16, device=g.device) if want_bf16 else None
    _lib.Stats.annotate(float(B * S * d) * ((4 if want_f32 else 0) + (2 if want_bf16 else 0)))
    _lib.call("sct_seq_mean_bwd", _ptr(g), _ptr(gx), _ptr(gy), B, S, d, _stream())
    return gx, gy


# ---------------------------------------------------------------------------------------------- K2
def _pick_bn(N, M=None, nt=False):
    # measured on B200 (tools/gemm_bench.py, M = 32768): the 128x256 tile wins from N = 768 up (880 vs 804
    # TFLOP/s at N = K = 768, 1045 vs 917 at N = 2304); the 128x128 tile only for the narrow N = 384 GEMMs.
    # One row tile (the decode step, M <= 128): the GEMM is a weight stream and N / 256 CTAs (3 for N = 768) leave the
    # machine empty -> 128x64 tiles (forward GEMM only)
    if nt and M is not None and M <= 128:
        return 64
    return 256 if N >= 512 else 128


def gemm_nt(a, w, bias=None, alpha=1.0, out=None, bn=None):
    """out[M,N] (bf16) = alpha * a[M,K] @ w[N,K]^T + bias"""
    a, lda = _rows2d(a, BF16, "a")
    w, ldw = _rows2d(w, BF16, "w")
    M, K = a.shape
    N, K2 = w.shape
    assert K == K2
    if out is None:  # row pitch padded to a multiple of 8 elements (TMA needs 16-byte pitches)
        out = torch.empty((M, (N + 7) // 8 * 8), dtype=BF16, device=a.device)[:, :N]
    out, ldd = _rows2d(out, BF16, "out")
    if bias is not None:
        _chk(bias, F32, "bias")
    _lib.Stats.annotate(2.0 * M * N * K)
    _lib.call("sct_gemm_bf16_nt", _ptr(a), lda, _ptr(w), ldw, _ptr(out), ldd, _ptr(bias), float(alpha),
              M, N, K, bn or _pick_bn(N, M, nt=True), _stream())
    return out


def gemm_nn(a, w, alpha=1.0, out=None, bn=None):
    """out[M,N] (bf16) = alpha * a[M,K] @ w[K,N]"""
    a, lda = _rows2d(a, BF16, "a")
    w, ldw = _rows2d(w, BF16, "w")
    M, K = a.shape
    K2, N = w.shape
    assert K == K2
    if out is None:
        out = torch.empty((M, (N + 7) // 8 * 8), dtype=BF16, device=a.device)[:, :N]
    out, ldd = _rows2d(out, BF16, "out")
    _lib.Stats.annotate(2.0 * M * N * K)
    _lib.call("sct_gemm_bf16_nn", _ptr(a), lda, _ptr(w), ldw, _ptr(out), ldd, None, float(alpha), M, N, K,
              bn or _pick_bn(N), _stream())
    return out


def gemm_nt_gelu(a, w, bias, p_drop=0.0, seed=0, offset=0, epoch=None):
    """(h, g): with z = a[M,K] @ w[N,K]^T + bias, h = dropout(gelu(z)) and g = mask / (1 - p) * gelu'(z) (bf16) —
    linear1 + activation + dropout of the feed-forward block in one kernel (fused GEMM epilogue); g is the local
    derivative the backward multiplies with (gemm_nn_mul)."""
    a, lda = _rows2d(a, BF16, "a")
    w, ldw = _rows2d(w, BF16, "w")
    M, K = a.shape
    N, K2 = w.shape
    assert K == K2 and N % 64 == 0
    h = torch.empty((M, N), dtype=BF16, device=a.device)
    g = torch.empty((M, N), dtype=BF16, device=a.device)
    if bias is not None:
        _chk(bias, F32, "bias")
    _lib.Stats.annotate(2.0 * M * N * K)
    _lib.call("sct_gemm_bf16_nt_gelu", _ptr(a), lda, _ptr(w), ldw, _ptr(h), N, _ptr(g), N, _ptr(bias), M, N, K,
              float(p_drop), seed, offset, _eptr(epoch), _stream())
    return h, g


def gemm_nn_mul(dy, w, g):
    """dz[M,N] (bf16) = (dy[M,K] @ w[K,N]) * g[M,N] — dgrad of linear2 + backward of the activation and its dropout in
    one kernel (g from gemm_nt_gelu)."""
    dy, lda = _rows2d(dy, BF16, "dy")
    w, ldw = _rows2d(w, BF16, "w")
    g, ldg = _rows2d(g, BF16, "g")
    M, K = dy.shape
    K2, N = w.shape
    assert K == K2 and g.shape == (M, N) and N % 64 == 0
    dz = torch.empty((M, N), dtype=BF16, device=dy.device)
    _lib.Stats.annotate(2.0 * M * N * K)
    _lib.call("sct_gemm_bf16_nn_mul", _ptr(dy), lda, _ptr(w), ldw, _ptr(g), ldg, _ptr(dz), N, M, N, K, _stream())
    return dz


def gemm_tn(a, b, out, alpha=1.0, k_splits=0, colsum=None):
    """out[M,N] (fp32) += alpha * a[K,M]^T @ b[K,N]; colsum[M] (fp32, optional) += alpha * a.sum(0)"""
    a, lda = _rows2d(a, BF16, "a")
    b, ldb = _rows2d(b, BF16, "b")
    K, M = a.shape
    K2, N = b.shape
    assert K == K2
    out, ldd = _rows2d(out, F32, "out")
    assert out.shape == (M, N)
    _lib.Stats.annotate(2.0 * M * N * K)
    if colsum is None:
        _lib.call("sct_gemm_bf16_tn", _ptr(a), lda, _ptr(b), ldb, _ptr(out), ldd, float(alpha), M, N, K,
                  k_splits, _stream())
    else:
        assert colsum.dtype == F32 and colsum.is_contiguous() and colsum.numel() >= M
        _lib.call("sct_gemm_bf16_tn_colsum", _ptr(a), lda, _ptr(b), ldb, _ptr(out), ldd, _sptr(colsum), float(alpha),
                  M, N, K, k_splits, _stream())
    return out


# ---------------------------------------------------------------------------------------------- K3
def attn_fwd(q, k, v, B, H, Lq, Lk, kpm=None, causal=False, scale=None, p_drop=0.0, seed=0, offset=0, epoch=None,
             head_dim=96, kv_batch_stride=0):
    """q [B*Lq, ld], k/v [B*Lk, ld] 2-D bf16 views (may be column slices of a packed projection).
    kv_batch_stride (elements) > 0: k/v are views into a [B, T_max, ld] cache of which rows [0, Lk) are used."""
    q, ldq = _rows2d(q, BF16, "q")
    k, ldk = _rows2d(k, BF16, "k")
    v, ldv = _rows2d(v, BF16, "v")
    assert ldk == ldv
    if scale is None:
        scale = head_dim ** -0.5
    o = torch.empty((B * Lq, H * head_dim), dtype=BF16, device=q.device)
    lse2 = torch.empty((B, H, Lq), dtype=F32, device=q.device)
    if kpm is not None:
        assert kpm.dtype in (torch.uint8, torch.bool) and kpm.is_contiguous() and kpm.shape == (B, Lk)
    _lib.Stats.annotate(4.0 * B * H * Lq * Lk * head_dim * (0.5 if causal else 1.0))
    _lib.call("sct_attn_fwd_strided", _ptr(q), ldq, _ptr(k), _ptr(v), ldk, int(kv_batch_stride), _ptr(o), o.stride(0),
              _ptr(lse2),
              kpm.data_ptr() if kpm is not None else None, B, H, Lq, Lk, head_dim, int(causal),
              float(scale), float(p_drop), seed, offset, _eptr(epoch), _stream())
    return o, lse2


def attn_bwd(q, k, v, o, d_o, lse2, B, H, Lq, Lk, dq, dk, dv, kpm=None, causal=False, scale=None,
             p_drop=0.0, seed=0, offset=0, epoch=None, head_dim=96, ws_slot=0):
    """dq [B*Lq, ld], dk/dv [B*Lk, ld] are written (2-D bf16 views, e.g. slices of a packed buffer)."""
    q, ldq = _rows2d(q, BF16, "q")
    k, ldk = _rows2d(k, BF16, "k")
    v, ldv = _rows2d(v, BF16, "v")
    assert ldk == ldv
    _chk(o, BF16, "o"), _chk(d_o, BF16, "d_o")
    dq, lddq = _rows2d(dq, BF16, "dq")
    dk, lddk = _rows2d(dk, BF16, "dk")
    dv, lddv = _rows2d(dv, BF16, "dv")
    assert lddk == lddv
    if scale is None:
        scale = head_dim ** -0.5
    dvec = torch.empty((B, H, Lq), dtype=F32, device=q.device)
    ws, ws_bytes = _attn_workspace(B, H, Lq, Lk, q.device, ws_slot)
    _lib.Stats.annotate(10.0 * B * H * Lq * Lk * head_dim * (0.5 if causal else 1.0))
    _lib.call("sct_attn_bwd_ws", _ptr(q), ldq, _ptr(k), _ptr(v), ldk, _ptr(o), _ptr(d_o), o.stride(0),
              _ptr(lse2), _ptr(dvec), _ptr(dq), lddq, _ptr(dk), _ptr(dv), lddk,
              kpm.data_ptr() if kpm is not None else None, B, H, Lq, Lk, head_dim, int(causal),
              float(scale), float(p_drop), seed, offset, _eptr(epoch), ws.data_ptr() if ws is not None else None, ws_bytes,
              _stream())


# dS^T workspace of the attention backward: one grow-only buffer per (device, slot); the calls of a step run back to
# back on one stream, work that runs concurrently on a side stream (the vulnerability heads) uses a slot of its own,
# chosen at forward time (`ws_slot`).  Outgrown buffers are kept alive because captured CUDA graphs may still point at
# them.
_ATTN_WS = {}
_WS_SLOT = 0


def current_ws_slot() -> int:
    return _WS_SLOT


class use_ws_slot:
    """`with use_ws_slot(1): ...` — attention ops recorded inside draw their backward workspace from slot 1."""

    def __init__(self, slot):
        self.slot = slot

    def __enter__(self):
        global _WS_SLOT
        self.prev, _WS_SLOT = _WS_SLOT, self.slot

    def __exit__(self, *exc):
        global _WS_SLOT
        _WS_SLOT = self.prev

_ATTN_WS_LIMIT = int(os.environ.get("SCT_ATTN_BWD_WS_MB", "16384")) << 20  # 0 disables (recomputing dQ kernel)


def _attn_workspace(B, H, Lq, Lk, device, slot=0):
    need = int(_lib.load().sct_attn_bwd_workspace_bytes(B, H, Lq, Lk))
    if need > _ATTN_WS_LIMIT:
        return None, 0
    ent = _ATTN_WS.setdefault((device, slot), [])
    if not ent or ent[-1].numel() < need:
        assert not torch.cuda.is_current_stream_capturing(), \
            "attention workspace must be sized by an eager step before CUDA-graph capture"
        ent.append(torch.empty(need, dtype=torch.uint8, device=device))
    return ent[-1], ent[-1].numel()


# --------------------------------------------------------------------------------------------- K4b
def ce_rows(logits, targets, V, grad_scale=1.0, write_grad=True):
    """logits [rows, ld>=V] bf16 (overwritten by the gradient when write_grad); returns (row_loss, row_lse)."""
    logits, ld = _rows2d(logits, BF16, "logits")
    rows = logits.shape[0]
    _chk(targets, torch.int64, "targets")
    row_loss = torch.empty(rows, dtype=F32, device=logits.device)
    row_lse = torch.empty(rows, dtype=F32, device=logits.device)
    _lib.Stats.annotate(float(rows) * V * (4 if write_grad else 2))  # read once (+ write the gradient once)
    _lib.call("sct_ce_rows", _ptr(logits), _ptr(targets), _ptr(row_loss), _ptr(row_lse), rows, V, ld,
              float(grad_scale), int(write_grad), _stream())
    return row_loss, row_lse


def sample_rows(logits, V, prev_tokens=None, temperature=0.7, top_k=50, top_p=0.95, greedy=False, seed=0, offset=0,
                epoch=None, out=None):
    """Next token per row of bf16 logits [rows, ld >= V]: temperature, syntax tweak (prev_tokens: int64 [rows] or None),
    top-k, top-p, multinomial — or argmax when greedy — in one launch (model.py:892-918).  Returns int64 [rows, 1]."""
    logits, ld = _rows2d(logits, BF16, "logits")
    rows = logits.shape[0]
    if prev_tokens is not None:
        _chk(prev_tokens, torch.int64, "prev_tokens")
        assert prev_tokens.numel() == rows
    if out is None:
        out = torch.empty((rows, 1), dtype=torch.int64, device=logits.device)
    assert out.dtype == torch.int64 and out.is_contiguous() and out.numel() == rows
    _lib.Stats.annotate(float(rows) * V * 2)  # the row is read once
    _lib.call("sct_sample_rows", _ptr(logits), ld, rows, V, float(temperature), int(top_k), float(top_p), int(bool(greedy)),
              _ptr(prev_tokens, 8), int(seed), int(offset), _eptr(epoch), _ptr(out, 8), _stream())
    return out


# ---------------------------------------------------------------------------------------------- K5
def small_linear_fwd(x, w, bias, out_bf16=False):
    M, K = x.shape
    N = w.shape[0]
    _chk(w, F32, "w")
    assert x.is_contiguous()
    xf = x if x.dtype == F32 else None
    xb = x if x.dtype == BF16 else None
    y = torch.empty((M, N), dtype=BF16 if out_bf16 else F32, device=x.device)
    _lib.Stats.annotate(float(N * K) * 4 + float(M * K) * x.element_size() + float(M * N) * y.element_size())
    _lib.call("sct_small_linear_fwd", _ptr(xf), _ptr(xb), _ptr(w), _ptr(bias), None if out_bf16 else _ptr(y),
              _ptr(y) if out_bf16 else None, M, N, K, _stream())
    return y


def small_linear_bwd(dy, x, w, need_dx=True, dx_bf16=False):
    M, K = x.shape
    N = w.shape[0]
    assert dy.is_contiguous() and x.is_contiguous()
    dyf = dy if dy.dtype == F32 else None
    dyb = dy if dy.dtype == BF16 else None
    xf = x if x.dtype == F32 else None
    xb = x if x.dtype == BF16 else None
    dw = torch.empty((N, K), dtype=F32, device=x.device)
    db = torch.empty((N,), dtype=F32, device=x.device)
    dx = torch.empty((M, K), dtype=BF16 if dx_bf16 else F32, device=x.device) if need_dx else None
    _lib.Stats.annotate(float(N * K) * 8 + float(M * K) * 2 * x.element_size() + float(M * N) * dy.element_size())
    _lib.call("sct_small_linear_bwd", _ptr(dyf), _ptr(dyb), _ptr(xf), _ptr(xb), _ptr(w),
              _ptr(dx) if (need_dx and not dx_bf16) else None, _ptr(dx) if (need_dx and dx_bf16) else None,
              _ptr(dw), _ptr(db), M, N, K, _stream())
    return dx, dw, db


def gan_loss_fwd(z, c_in=None):
    _chk(z, F32, "z")
    out4 = torch.empty(4, dtype=F32, device=z.device)
    _lib.call("sct_gan_loss_fwd", _sptr(z), z.numel(), _sptr(c_in), _sptr(out4), _stream())
    return out4


def gan_loss_bwd(z, c, g_d, g_adv):
    dz = torch.empty_like(z)
    _lib.call("sct_gan_loss_bwd", _sptr(z), z.numel(), _sptr(c), _sptr(g_d), _sptr(g_adv), _sptr(dz), _stream())
    return dz


# ------------------------------------------------------------------------------------- optimiser tail
def clip_adamw_step(table, n_tensors, chunks, n_chunks, loss, sqnorm3, out2, max_norm, disc_mult, vuln_mult,
                    beta1, beta2, eps, n_elems=0):
    """table: uint8 device tensor of n_tensors sct_opt_tensor records; chunks: int32 [n_chunks, 2]."""
    # norm pass reads g; update pass reads g, p, m, v and writes p, m, v (+ the bf16 weight copy): 34 B / parameter
    _lib.Stats.annotate(float(n_elems) * 34)
    _lib.call("sct_clip_adamw_step", _ptr(table), n_tensors, _ptr(chunks, 8), n_chunks, _sptr(loss), _sptr(sqnorm3),
              _sptr(out2), float(max_norm), float(disc_mult), float(vuln_mult), float(beta1), float(beta2),
              float(eps), _stream())
