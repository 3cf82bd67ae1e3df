"""Autograd layer over the C-ABI kernels (kernels.py): one torch.autograd.Function per fused op.

Conventions
  * the residual stream is fp32 [rows, d]; everything fed to a GEMM is bf16 [rows, *]
  * master parameters stay fp32 nn.Parameters (state_dict-compatible with the reference); GEMMs read a
    private bf16 shadow (ShadowCache) that is refreshed when the parameter's version counter moves
  * dropout masks are never stored: every site draws a (seed, offset) pair from DropoutRng in forward
    and replays it in backward
  * weight gradients are produced in fp32 and returned to autograd for the fp32 parameter, so the
    reference's parameter hooks (feature_fusion clamp, model.py:285-286) and optimizer groups work
    unchanged
"""
from __future__ import annotations

import torch

from . import kernels as kn

BF16, F32 = torch.bfloat16, torch.float32


# ------------------------------------------------------------------------------------------------
class DropoutRng:
    """(seed, offset) source for the counter-based dropout of the kernels.

    `seed` is fixed per model (from torch's seed), `offset` numbers the dropout sites of one forward pass
    (reset by begin_step), and a device-resident epoch counter owned by the MODEL — bumped once per training
    forward ON THE DEVICE, so that it also advances when the step is replayed from a CUDA graph — is folded in
    by the kernels at run time.  The counter's tensor is handed to every kernel call (the C ABI's `epoch`
    argument) and kept in the op's autograd context, so a backward always replays the masks of its own
    forward even if another model ran a forward in between; the library holds no dropout state.  Per model:
    the backward must run before that model's next training forward bumps its epoch."""
    seed = 0x5C7B200
    counter = 0
    epoch = None

    @classmethod
    def reseed(cls, seed: int):
        cls.seed = int(seed) & 0x7FFFFFFFFFFFFFFF
        cls.counter = 0

    @classmethod
    def begin_step(cls, owner, device):
        """Start a training forward of `owner` (the module): site counter to 0, its device epoch += 1.  The
        epoch counter and the seed (torch.initial_seed() when the module first trains) belong to the module,
        so a freshly built model under the same torch seed reproduces the same masks."""
        ep = getattr(owner, "_drop_epoch", None)
        if ep is None or ep.device != device:
            ep = torch.zeros(2, dtype=torch.int64, device=device)  # [0] = epoch (16-byte allocation)
            owner._drop_epoch = ep
            rank = 0
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                rank = torch.distributed.get_rank()  # replicas see different samples: give them different masks too
            owner._drop_seed = (torch.initial_seed() * 1000003 + rank * 7919 + 0x5C7B200) & 0x7FFFFFFFFFFFFFFF
        cls.seed = owner._drop_seed
        cls.epoch = ep
        ep[:1].add_(1)
        cls.counter = 0

    @classmethod
    def attach(cls, owner, device):
        """Draw the following sites from `owner`'s stream WITHOUT advancing its epoch (sub-module shims): the site
        numbering continues from a per-model counter kept apart from the numbers `forward` uses."""
        if getattr(owner, "_drop_epoch", None) is None or owner._drop_epoch.device != device:
            cls.begin_step(owner, device)  # first use: creates the epoch tensor and the seed
        cls.seed, cls.epoch = owner._drop_seed, owner._drop_epoch
        nxt = getattr(owner, "_shim_sites", 0x8000)
        cls.counter = nxt if nxt < 0xF0000 else 0x8000

    @classmethod
    def detach_sites(cls, owner):
        owner._shim_sites = cls.counter

    @classmethod
    def draw(cls):
        cls.counter += 1
        return cls.seed, cls.counter, cls.epoch


class ShadowCache:
    """bf16 copies of fp32 parameters for the tensor-core GEMMs (private, non-persistent).

    Fused optimisers (torch's and csrc/optim.cu) update parameters without bumping Tensor._version, so a training
    forward cannot tell from the version whether a copy is stale.  Two ways a copy is known to be current:
      * the library's own optimiser kernel rewrote it together with the weight (`mark_synced`, after every
        `sct_clip_adamw_step` that was given the copy's address): the next forward uses it as is — no cast pass;
      * it was cast earlier in the same pass.
    Otherwise a training forward re-casts the weight on first use (once per pass), and inference re-casts a copy made
    during training on first use and then keeps it until the parameter's version or storage changes
    (load_state_dict, .to()).  Buffers are reused, so their addresses are stable across casts (captured CUDA graphs and
    the optimiser's pointer table keep referring to the right memory)."""

    def __init__(self):
        self._store = {}  # id(param) -> [version, bf16 buffer, data_ptr, reusable_in_eval, synced_by_optimiser]
        self.training = False
        self._fresh = set()

    def begin_step(self, refresh: bool):
        self.training = refresh
        self._fresh.clear()

    def _entry(self, p):
        ent = self._store.get(id(p))
        if ent is not None and ent[1].device == p.device and ent[2] == p.data_ptr() and ent[1].shape == p.shape:
            return ent
        return None

    def get(self, p: torch.Tensor) -> torch.Tensor:
        key = id(p)
        ent = self._entry(p)
        if ent is not None:
            if ent[4] and ent[0] == p._version:
                return ent[1]
            if self.training:
                if key in self._fresh:
                    return ent[1]
            elif ent[3] and ent[0] == p._version:
                return ent[1]
        src = p.detach()
        src2 = src if src.dim() == 2 else src.view(1, -1)
        assert src2.is_contiguous()
        buf = ent[1] if ent is not None else torch.empty(src.shape, dtype=BF16, device=p.device)
        kn.cast_scale(src2, buf.view(src2.shape), 0, 1.0)
        self._store[key] = [p._version, buf, p.data_ptr(), not self.training, False]
        self._fresh.add(key)
        return buf

    def peek_ptr(self, p) -> int:
        """Address of p's bf16 copy if it has one (for the optimiser's pointer table), else 0."""
        ent = self._entry(p)
        return ent[1].data_ptr() if ent is not None else 0

    def mark_synced(self, params):
        """The optimiser kernel has just rewritten the copies of `params` from the updated weights."""
        for p in params:
            ent = self._entry(p)
            if ent is not None:
                ent[0], ent[4] = p._version, True

    def clear(self):
        self._store.clear()


# ------------------------------------------------------------------------------------------------
class GradArena:
    """Zero-initialised fp32 gradient buffers for one training step, carved from ONE arena that is cleared by ONE
    memset at the start of the step.

    Every parameter gradient of the step is an accumulator (split-K wgrad GEMMs reduce-add into it, the embedding
    backward scatter-adds, LayerNorm gamma / beta gradients are atomics), so each used to be a `torch.zeros` of its
    own: 300+ fill kernels per step.  The trainer brackets a step with `begin()` / `end()`; autograd backward
    functions call `zeros()` and get a view of the arena (same addresses every step: the optimiser's pointer table
    stays valid, eager or captured).  Outside a bracket — or while the arena is still being sized by the first
    step — `zeros()` is plain `torch.zeros`.  A parameter used at several sites of one step (the shared
    `embedding.weight`: contract and target passes) gets ONE buffer: `zeros_for(param)` returns it again and the
    second backward accumulates into it instead of producing a second 154 MB tensor for autograd to add."""

    ALIGN = 64  # elements (256 bytes)

    def __init__(self):
        self.buf = None
        self.off = 0
        self.need = 0       # elements requested so far in the running step
        self.need_last = 0  # ... in the largest completed step: what begin() sizes the arena for
        self.active = False
        self.shared = {}
        self._old = []      # outgrown arenas stay alive: captured graphs may still point into them

    def reserve(self, device):
        """Allocate (outside any graph capture) what the last completed step asked for."""
        if self.need_last > 0 and (self.buf is None or self.buf.device != device or self.need_last > self.buf.numel()):
            assert not (device.type == "cuda" and torch.cuda.is_current_stream_capturing()), \
                "gradient arena must be sized by an eager step before CUDA-graph capture"
            if self.buf is not None:
                self._old.append(self.buf)
            self.buf = torch.empty(self.need_last, dtype=F32, device=device)

    def begin(self, device):
        if not (device.type == "cuda" and torch.cuda.is_current_stream_capturing()):
            self.reserve(device)
        self.off, self.need, self.shared = 0, 0, {}
        self.active = True
        if self.buf is not None:
            self.buf.zero_()

    def end(self):
        self.active = False
        self.shared = {}
        self.need_last = max(self.need_last, self.need)

    def zeros(self, shape, device):
        if not self.active:
            return torch.zeros(shape, dtype=F32, device=device)
        n = 1
        for s in shape:
            n *= int(s)
        n_al = (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.need += n_al
        if self.buf is None or self.buf.device != device or self.off + n_al > self.buf.numel():
            return torch.zeros(shape, dtype=F32, device=device)  # sizing pass (first step) / outgrown
        out = self.buf[self.off:self.off + n].view(shape)
        self.off += n_al
        return out

    def zeros_for(self, param, shape, device):
        """(buffer, first): the accumulator of `param` for this step; `first` is False when an earlier backward of
        the same step already returned it to autograd (return None for the parameter then)."""
        if not self.active:
            return torch.zeros(shape, dtype=F32, device=device), True
        key = param if isinstance(param, int) else param.data_ptr()  # saved_tensors hands out fresh wrappers: key by address
        ent = self.shared.get(key)
        if ent is not None:
            return ent, False
        buf = self.zeros(shape, device)
        # keep an ALIAS, hand out the original: autograd's AccumulateGrad only adopts a gradient it holds the sole
        # reference to; a tensor that is also referenced from here would be cloned (a 154 MB copy per table and step)
        self.shared[key] = buf.detach()
        return buf, True


ARENA = GradArena()


def grad_zeros(shape, device):
    return ARENA.zeros(tuple(shape), device)


class EmbedLnPe(torch.autograd.Function):
    """K1: LayerNorm(dropout(table[ids] * sqrt(d))) + pe[s]  (model.py:412-421, 944-947)."""

    @staticmethod
    def forward(ctx, ids, table, gamma, beta, pe, seq_len, scale, p_drop, want_f32, want_bf16):
        seed, off, ep = DropoutRng.draw() if p_drop > 0 else (0, 0, None)
        ids_flat = ids.reshape(-1).contiguous()
        out_f32, out_bf16, stats = kn.embed_ln_pe_fwd(ids_flat, table.detach(), gamma.detach(), beta.detach(),
                                                      pe, seq_len, scale, p_drop, seed, off, ep, want_f32, want_bf16)
        ctx.save_for_backward(ids_flat, table, gamma, stats)
        ctx.beta_key = beta.data_ptr()
        ctx.cfg = (scale, p_drop, seed, off, ep)
        ctx.set_materialize_grads(False)
        return out_f32, out_bf16

    @staticmethod
    def backward(ctx, g_f32, g_bf16):
        ids_flat, table, gamma, stats = ctx.saved_tensors
        scale, p_drop, seed, off, ep = ctx.cfg
        g, g2 = (g_f32, g_bf16) if g_f32 is not None else (g_bf16, None)
        # the table / gamma / beta accumulators are shared by every pass that uses this embedding in the step
        dtable, first = ARENA.zeros_for(table, table.shape, table.device)
        dgamma, _ = ARENA.zeros_for(gamma, gamma.shape, gamma.device)
        dbeta, _ = ARENA.zeros_for(ctx.beta_key, gamma.shape, gamma.device)
        if g is not None:
            kn.embed_ln_pe_bwd(g.contiguous(), ids_flat, table.detach(), gamma.detach(), stats, dtable, dgamma,
                               dbeta, scale, p_drop, seed, off, ep, g2=None if g2 is None else g2.contiguous())
        if not first:  # already handed to autograd by the other pass: accumulated in place
            return None, None, None, None, None, None, None, None, None, None
        return None, dtable, dgamma, dbeta, None, None, None, None, None, None


def embed_ln_pe(ids, table, gamma, beta, pe, seq_len, scale, p_drop=0.0, want_f32=True, want_bf16=False):
    return EmbedLnPe.apply(ids, table, gamma, beta, pe, seq_len, scale, p_drop, want_f32, want_bf16)


# ------------------------------------------------------------------------------------------------
class ResidualLn(torch.autograd.Function):
    """K4a: x' = x + alpha * dropout(branch); second output = LayerNorm(x') or bf16(x').

    mode: 'ln' (needs gamma/beta), 'cast', or 'none'.  Returns (x' fp32 or None, second bf16 or None)."""

    @staticmethod
    def forward(ctx, x, branch, gamma, beta, alpha, p_drop, mode, want_x):
        use_drop = p_drop > 0 and branch is not None
        seed, off, ep = DropoutRng.draw() if use_drop else (0, 0, None)
        p = p_drop if use_drop else 0.0
        g_ = gamma.detach() if gamma is not None else None
        b_ = beta.detach() if beta is not None else None
        need_xprime = want_x or mode == "ln"
        x_out, y_ln, y_cast, stats = kn.add_dropout_ln_fwd(
            x, branch, alpha, g_, b_, want_x=need_xprime, want_ln=(mode == "ln"), want_cast=(mode == "cast"),
            p_drop=p, seed=seed, offset=off, epoch=ep)
        ctx.mode, ctx.alpha, ctx.drop = mode, alpha, (p, seed, off, ep)
        ctx.has_x, ctx.has_branch = x is not None, branch is not None
        if mode == "ln":
            ctx.save_for_backward(x_out, stats, gamma)
        ctx.set_materialize_grads(False)
        return (x_out if want_x else None), (y_ln if mode == "ln" else y_cast)

    @staticmethod
    def backward(ctx, g_xout, g_second):
        p, seed, off, ep = ctx.drop
        dgamma = dbeta = None
        g_yln = g_ycast = None
        xprime = stats = gamma = None
        if ctx.mode == "ln":
            xprime, stats, gamma = ctx.saved_tensors
            if g_second is not None:
                g_yln = g_second.contiguous()
            dgamma = grad_zeros(gamma.shape, gamma.device)
            dbeta = grad_zeros(gamma.shape, gamma.device)
        elif ctx.mode == "cast" and g_second is not None:
            g_ycast = g_second.contiguous()
        if g_xout is None and g_yln is None and g_ycast is None:
            return None, None, dgamma, dbeta, None, None, None, None
        if g_xout is not None:
            g_xout = g_xout.contiguous()
        g_x, g_branch = kn.add_dropout_ln_bwd(
            g_xout, g_yln, g_ycast, xprime if g_yln is not None else None, stats if g_yln is not None else None,
            gamma.detach() if g_yln is not None else None, ctx.alpha, dgamma if g_yln is not None else None,
            dbeta if g_yln is not None else None, want_gx=ctx.has_x, want_gbranch=ctx.has_branch,
            p_drop=p, seed=seed, offset=off, epoch=ep)
        return g_x, g_branch, dgamma, dbeta, None, None, None, None


def residual_ln(x, branch, gamma=None, beta=None, alpha=1.0, p_drop=0.0, mode="ln", want_x=True):
    return ResidualLn.apply(x, branch, gamma, beta, alpha, p_drop, mode, want_x)


class LnAct(torch.autograd.Function):
    """dropout(gelu(LayerNorm(z))) on bf16 rows (model.py:225-235, 253-271)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, p_drop):
        seed, off, ep = DropoutRng.draw() if p_drop > 0 else (0, 0, None)
        h, stats = kn.ln_act_fwd(z, gamma.detach(), beta.detach(), p_drop, seed, off, ep)
        ctx.save_for_backward(z, stats, gamma, beta)
        ctx.drop = (p_drop, seed, off, ep)
        return h

    @staticmethod
    def backward(ctx, g_h):
        z, stats, gamma, beta = ctx.saved_tensors
        p, seed, off, ep = ctx.drop
        dgamma, dbeta = grad_zeros(gamma.shape, gamma.device), grad_zeros(beta.shape, beta.device)
        g_z = kn.ln_act_bwd(g_h.contiguous(), z, stats, gamma.detach(), beta.detach(), dgamma, dbeta, p, seed, off, ep)
        return g_z, dgamma, dbeta, None


def ln_act(z, gamma, beta, p_drop=0.0):
    return LnAct.apply(z, gamma, beta, p_drop)


class GeluDropout(torch.autograd.Function):
    """dropout(gelu_erf(z)) — the FFN activation (torch transformer.py _ff_block)."""

    @staticmethod
    def forward(ctx, z, p_drop):
        seed, off, ep = DropoutRng.draw() if p_drop > 0 else (0, 0, None)
        h = kn.gelu_dropout_fwd(z, p_drop, seed, off, ep)
        ctx.save_for_backward(z)
        ctx.drop = (p_drop, seed, off, ep)
        return h

    @staticmethod
    def backward(ctx, g_h):
        (z,) = ctx.saved_tensors
        p, seed, off, ep = ctx.drop
        return kn.gelu_dropout_bwd(g_h.contiguous(), z, p, seed, off, ep), None


def gelu_dropout(z, p_drop=0.0):
    return GeluDropout.apply(z, p_drop)


class ConcatScaled(torch.autograd.Function):
    """[a, scale_b * b] along features without torch.cat (model.py:450)."""

    @staticmethod
    def forward(ctx, a, b, scale_b):
        rows, da = a.shape
        db = b.shape[1]
        out = torch.empty((rows, da + db), dtype=BF16, device=a.device)
        kn.cast_scale(a, out, 0, 1.0)
        kn.cast_scale(b, out, da, scale_b)
        ctx.dims = (da, db, scale_b)
        return out

    @staticmethod
    def backward(ctx, g):
        da, db, scale_b = ctx.dims
        g = g.contiguous()
        ga = torch.empty((g.shape[0], da), dtype=BF16, device=g.device)
        gb = torch.empty((g.shape[0], db), dtype=BF16, device=g.device)
        kn.cast_scale(g[:, :da], ga, 0, 1.0)
        kn.cast_scale(g[:, da:], gb, 0, scale_b)
        return ga, gb, None


def concat_scaled(a, b, scale_b):
    return ConcatScaled.apply(a, b, scale_b)


# ------------------------------------------------------------------------------------------------
class Linear(torch.autograd.Function):
    """K2: y = x W^T + b on tcgen05 (forward nt, dgrad nn, wgrad tn split-K fp32, bias grad colsum)."""

    @staticmethod
    def forward(ctx, x, w_f32, bias, w_bf16):
        y = kn.gemm_nt(x, w_bf16, bias.detach() if bias is not None else None)
        ctx.save_for_backward(x, w_bf16)
        ctx.has_bias = bias is not None
        ctx.w_shape = w_f32.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w_bf16 = ctx.saved_tensors
        N = ctx.w_shape[0]
        ld = (N + 7) // 8 * 8
        if ld != N:  # ragged vocab: TMA wants 16-byte row pitches -> zero-padded copy
            full = torch.zeros((dy.shape[0], ld), dtype=BF16, device=dy.device)
            full[:, :N] = dy
            dy = full[:, :N]
        else:
            full = dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = kn.gemm_nn(dy, w_bf16)
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        if want_db:
            db = grad_zeros((ld,), dy.device)
        if ctx.needs_input_grad[1]:
            dw = grad_zeros(ctx.w_shape, dy.device)
            kn.gemm_tn(dy, x, dw, colsum=db)  # bias gradient summed from the dY tiles of the wgrad GEMM
        elif want_db:
            kn.colsum_bf16(full, db)
        if want_db:
            db = db[:N]
        return dx, dw, db, None


def linear(x, w_f32, bias, w_bf16):
    return Linear.apply(x, w_f32, bias, w_bf16)


class FusedFFN(torch.autograd.Function):
    """The feed-forward block linear2(dropout(gelu(linear1(y)))) (torch transformer.py _ff_block, model.py:56-77) as
    GEMM launches only: the activation and its dropout run in linear1's epilogue, which also leaves their local
    derivative g = mask / (1 - p) * gelu'(z) for the backward; linear2's dgrad multiplies with it on the way out — no
    separate pass over the [rows, ff] activations in either direction."""

    @staticmethod
    def forward(ctx, y, w1_f32, b1, w1_bf16, w2_f32, b2, w2_bf16, p_drop):
        seed, off, ep = DropoutRng.draw() if p_drop > 0 else (0, 0, None)
        h, g = kn.gemm_nt_gelu(y, w1_bf16, b1.detach(), p_drop, seed, off, ep)
        out = kn.gemm_nt(h, w2_bf16, b2.detach())
        ctx.save_for_backward(y, g, h, w1_bf16, w2_bf16)
        ctx.shapes = (w1_f32.shape, w2_f32.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, g, h, w1_bf16, w2_bf16 = ctx.saved_tensors
        s1, s2 = ctx.shapes
        dout = dout.contiguous()
        dev = dout.device
        dz = kn.gemm_nn_mul(dout, w2_bf16, g)
        dw2, db2 = grad_zeros(s2, dev), grad_zeros((s2[0],), dev)
        kn.gemm_tn(dout, h, dw2, colsum=db2)
        dy = kn.gemm_nn(dz, w1_bf16) if ctx.needs_input_grad[0] else None
        dw1, db1 = grad_zeros(s1, dev), grad_zeros((s1[0],), dev)
        kn.gemm_tn(dz, y, dw1, colsum=db1)
        return dy, dw1, db1, None, dw2, db2, None, None


def fused_ffn(y, lin1, lin2, w1_bf16, w2_bf16, p_drop=0.0):
    return FusedFFN.apply(y, lin1.weight, lin1.bias, w1_bf16, lin2.weight, lin2.bias, w2_bf16, p_drop)


class CrossProj(torch.autograd.Function):
    """The packed in-projection of a cross-attention (torch `_in_projection_packed`, functional.py:5798): q from rows
    [0, d) of in_proj_weight applied to the query stream, k|v from rows [d, 3d) applied to the memory / path stream.
    One autograd node for both, so the weight gradient is ONE [3d, d] buffer the two wgrad GEMMs write their row
    ranges of (autograd on `weight[r0:r1]` slices pads each slice gradient into a zero [3d, d] tensor and adds)."""

    @staticmethod
    def forward(ctx, xq, xkv, w_f32, bias, w_bf16):
        d = w_bf16.shape[1]
        b = bias.detach()
        q = kn.gemm_nt(xq, w_bf16[:d], b[:d])
        kv = kn.gemm_nt(xkv, w_bf16[d:], b[d:])
        ctx.save_for_backward(xq, xkv, w_bf16)
        return q, kv

    @staticmethod
    def backward(ctx, dq, dkv):
        xq, xkv, w_bf16 = ctx.saved_tensors
        d = w_bf16.shape[1]
        dq, dkv = dq.contiguous(), dkv.contiguous()
        dxq = kn.gemm_nn(dq, w_bf16[:d]) if ctx.needs_input_grad[0] else None
        dxkv = kn.gemm_nn(dkv, w_bf16[d:]) if ctx.needs_input_grad[1] else None
        dw = grad_zeros(w_bf16.shape, dq.device)
        db = grad_zeros((3 * d,), dq.device)
        kn.gemm_tn(dq, xq, dw[:d], colsum=db[:d])
        kn.gemm_tn(dkv, xkv, dw[d:], colsum=db[d:])
        return dxq, dxkv, dw, db, None


def cross_proj(xq, xkv, w_f32, bias, w_bf16):
    return CrossProj.apply(xq, xkv, w_f32, bias, w_bf16)


# ------------------------------------------------------------------------------------------------
class SelfAttention(torch.autograd.Function):
    """K3 on a packed [B*L, 3d] q|k|v projection."""

    @staticmethod
    def forward(ctx, qkv, B, H, L, kpm, causal, p_drop):
        d = qkv.shape[1] // 3
        seed, off, ep = DropoutRng.draw() if p_drop > 0 else (0, 0, None)
        o, lse2 = kn.attn_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, H, L, L, kpm=kpm, causal=causal,
                              p_drop=p_drop, seed=seed, offset=off, epoch=ep, head_dim=d // H)
        ctx.save_for_backward(qkv, o, lse2, kpm if kpm is not None else torch.empty(0))
        ctx.cfg = (B, H, L, causal, p_drop, seed, off, kpm is not None, ep)
        ctx.ws_slot = kn.current_ws_slot()
        return o

    @staticmethod
    def backward(ctx, d_o):
        qkv, o, lse2, kpm = ctx.saved_tensors
        B, H, L, causal, p_drop, seed, off, has_kpm, ep = ctx.cfg
        d = qkv.shape[1] // 3
        dqkv = torch.empty_like(qkv)
        kn.attn_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], o, d_o.contiguous(), lse2, B, H, L, L,
                    dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], kpm=kpm if has_kpm else None, causal=causal,
                    p_drop=p_drop, seed=seed, offset=off, epoch=ep, head_dim=d // H, ws_slot=ctx.ws_slot)
        return dqkv, None, None, None, None, None, None


class CrossAttention(torch.autograd.Function):
    """K3 with q [B*Lq, d] and a packed [B*Lk, 2d] k|v projection of the memory / path stream."""

    @staticmethod
    def forward(ctx, q, kv, B, H, Lq, Lk, kpm, p_drop):
        d = q.shape[1]
        seed, off, ep = DropoutRng.draw() if p_drop > 0 else (0, 0, None)
        o, lse2 = kn.attn_fwd(q, kv[:, :d], kv[:, d:], B, H, Lq, Lk, kpm=kpm, causal=False, p_drop=p_drop,
                              seed=seed, offset=off, epoch=ep, head_dim=d // H)
        ctx.save_for_backward(q, kv, o, lse2, kpm if kpm is not None else torch.empty(0))
        ctx.cfg = (B, H, Lq, Lk, p_drop, seed, off, kpm is not None, ep)
        ctx.ws_slot = kn.current_ws_slot()
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, kv, o, lse2, kpm = ctx.saved_tensors
        B, H, Lq, Lk, p_drop, seed, off, has_kpm, ep = ctx.cfg
        d = q.shape[1]
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        kn.attn_bwd(q, kv[:, :d], kv[:, d:], o, d_o.contiguous(), lse2, B, H, Lq, Lk, dq, dkv[:, :d], dkv[:, d:],
                    kpm=kpm if has_kpm else None, causal=False, p_drop=p_drop, seed=seed, offset=off,
                    epoch=ep, head_dim=d // H, ws_slot=ctx.ws_slot)
        return dq, dkv, None, None, None, None, None, None


def self_attention(qkv, B, H, L, kpm=None, causal=False, p_drop=0.0):
    return SelfAttention.apply(qkv, B, H, L, kpm, causal, p_drop)


def cross_attention(q, kv, B, H, Lq, Lk, kpm=None, p_drop=0.0):
    return CrossAttention.apply(q, kv, B, H, Lq, Lk, kpm, p_drop)


@torch.no_grad()
def cached_attention(q, kv_cache, B, H, Lk, kpm=None):
    """Decode-step attention (inference only): q [B, d] (one new position per sequence) against rows [0, Lk) of a
    [B, T_max, 2d] k|v cache; no dropout, no autograd."""
    d = q.shape[1]
    t_max = kv_cache.shape[1]
    flat = kv_cache.view(B * t_max, 2 * d)
    o, _ = kn.attn_fwd(q, flat[:, :d], flat[:, d:], B, H, 1, Lk, kpm=kpm, causal=False, head_dim=d // H,
                       kv_batch_stride=t_max * 2 * d)
    return o


# ------------------------------------------------------------------------------------------------
class VocabCE(torch.autograd.Function):
    """K4b: mean token cross-entropy of (h W^T + b) against targets without materialising [rows, V].

    Rows are processed in chunks: vocab GEMM into a bf16 scratch chunk -> sct_ce_rows (loss + in-place
    d logits) -> dgrad / wgrad GEMMs.  Gradients w.r.t. h, W, b are therefore produced in forward (scaled
    by 1/n_valid) and only multiplied by the upstream scalar in backward.  targets < 0 exclude a row
    (the t = T-1 position of each sequence, model.py:962-964)."""

    @staticmethod
    def forward(ctx, h, w_f32, bias, w_bf16, targets, n_valid, chunk_rows):
        rows, d = h.shape
        V = w_bf16.shape[0]
        ld = (V + 7) // 8 * 8
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        dev = h.device
        chunk_rows = min(chunk_rows, rows)
        scratch = torch.empty((chunk_rows, ld), dtype=BF16, device=dev)
        if ld != V:
            scratch[:, V:].zero_()
        row_loss = torch.empty(rows, dtype=F32, device=dev)
        row_lse = torch.empty(rows, dtype=F32, device=dev)
        dh = torch.empty_like(h) if need_grad else None
        dw = grad_zeros((V, d), dev) if need_grad else None
        db = grad_zeros((V,), dev) if (need_grad and bias is not None) else None
        inv = 1.0 / float(n_valid)
        b_ = bias.detach() if bias is not None else None
        for r0 in range(0, rows, chunk_rows):
            r1 = min(r0 + chunk_rows, rows)
            n = r1 - r0
            logits = scratch[:n, :V]
            kn.gemm_nt(h[r0:r1], w_bf16, b_, out=logits)
            rl, rs = kn.ce_rows(logits, targets[r0:r1], V, grad_scale=inv, write_grad=need_grad)
            row_loss[r0:r1] = rl
            row_lse[r0:r1] = rs
            if need_grad:
                kn.gemm_nn(logits, w_bf16, out=dh[r0:r1])
                kn.gemm_tn(logits, h[r0:r1], dw, colsum=db)  # + bias gradient from the same tiles
        loss = row_loss.sum() * inv
        if need_grad:
            ctx.save_for_backward(dh, dw, db if db is not None else torch.empty(0))
        ctx.has_bias = db is not None
        ctx.mark_non_differentiable(row_lse)
        return loss, row_lse

    @staticmethod
    def backward(ctx, g_loss, _g_lse):
        dh, dw, db = ctx.saved_tensors
        g = g_loss.to(F32)
        # in place: these tensors were produced by this op's forward for exactly this use
        dh.mul_(g.to(BF16))
        dw.mul_(g)
        if ctx.has_bias:
            db.mul_(g)
        return dh, dw, db if ctx.has_bias else None, None, None, None, None


def _vocab_chunk_rows(rows, d, V, device):
    """Rows per chunk of the vocab projection.  The dgrad GEMM of a chunk (dH = dLogits W: K = V, only d / 256 column
    tiles) has ceil(chunk / 256) * ceil(d / 256) work items for the SMs / 2 CTA pairs: 4096-row chunks fill 48 of 74
    pairs for a full wave each (8 waves for 32768 rows), 5464-row chunks fill 66 of 74 (6 waves).  Pick the chunk
    count with the fewest dgrad waves overall, the scratch capped at 1 GiB; ties go to fewer chunks (each one
    reduce-adds the whole [V, d] weight gradient)."""
    units = max(1, torch.cuda.get_device_properties(device).multi_processor_count // 2)
    ld = (V + 7) // 8 * 8
    n_tiles = (d + 255) // 256
    best = None
    for c in range(1, 129):
        cr = ((rows + c - 1) // c + 7) // 8 * 8
        if cr * ld * 2 > (1 << 30) and c < 128:
            continue
        waves = ((cr + 255) // 256 * n_tiles + units - 1) // units
        cost = (waves * c, c)
        if best is None or cost < best[0]:
            best = (cost, cr)
    return best[1]


def vocab_ce(h, w_f32, bias, w_bf16, targets, n_valid, chunk_rows=None):
    if chunk_rows is None:
        chunk_rows = _vocab_chunk_rows(h.shape[0], h.shape[1], w_bf16.shape[0], h.device)
    return VocabCE.apply(h, w_f32, bias, w_bf16, targets, n_valid, chunk_rows)


# ------------------------------------------------------------------------------------------------
class SeqMean(torch.autograd.Function):
    """mean over the sequence of (x_f32 + y_bf16) (model.py:1187-1193 pooled before the linear projection)."""

    @staticmethod
    def forward(ctx, x, y, B, S):
        d = (x if x is not None else y).shape[1]
        ctx.cfg = (B, S, d, x is not None, y is not None)
        return kn.seq_mean_fwd(x, y, B, S, d)

    @staticmethod
    def backward(ctx, g):
        B, S, d, has_x, has_y = ctx.cfg
        gx, gy = kn.seq_mean_bwd(g.contiguous(), B, S, d, want_f32=has_x, want_bf16=has_y)
        return gx, gy, None, None


def seq_mean(x, y, B, S):
    return SeqMean.apply(x, y, B, S)


class SmallLinear(torch.autograd.Function):
    """K5: fp32 nn.Linear on [batch, K] rows (discriminator head MLPs, model.py:250-271)."""

    @staticmethod
    def forward(ctx, x, w, bias, out_bf16):
        y = kn.small_linear_fwd(x.contiguous(), w.detach(), bias.detach() if bias is not None else None, out_bf16)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dw, db = kn.small_linear_bwd(dy.contiguous(), x.contiguous(), w.detach(), need_dx=True,
                                         dx_bf16=(x.dtype == BF16))
        return dx, dw, (db if ctx.has_bias else None), None


def small_linear(x, w, bias, out_bf16=False):
    return SmallLinear.apply(x, w, bias, out_bf16)


class GanLoss(torch.autograd.Function):
    """K5: (d_loss, adv_loss, confidence) of train.py:1201-1234 with device-side 0.3 / 0.8 predicates."""

    @staticmethod
    def forward(ctx, z, c_in):
        zf = z.reshape(-1).contiguous()
        out4 = kn.gan_loss_fwd(zf, c_in)
        ctx.save_for_backward(zf, out4)
        ctx.shape = z.shape
        conf = out4[2].clone()
        ctx.mark_non_differentiable(conf)
        return out4[0].clone(), out4[1].clone(), conf

    @staticmethod
    def backward(ctx, g_d, g_adv, _g_c):
        zf, out4 = ctx.saved_tensors
        z0 = torch.zeros(1, dtype=F32, device=zf.device)
        gd = g_d.reshape(1).to(F32).contiguous() if g_d is not None else z0
        ga = g_adv.reshape(1).to(F32).contiguous() if g_adv is not None else z0
        dz = kn.gan_loss_bwd(zf, out4[2:3], gd, ga)
        return dz.view(ctx.shape), None


def gan_loss(z, c_in=None):
    return GanLoss.apply(z, c_in)
