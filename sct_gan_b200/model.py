"""B200-native drop-in for the reference's `SmartContractTransformer` (SCT-GAN/model.py:23-1217).

Same constructor, same `forward` kwargs, same returned dict keys, same sub-module attribute names and the
same 312 `state_dict` keys (fp32 nn.Parameters, `pos_encoder.pe` buffer, `path_embedding` alias), so the
reference's train.py / inference.py call sites and checkpoints work unchanged.  The arithmetic of the hot
path — dual embedding + positional encoding (K1), encoder / AST attention / feature fusion / decoder
(K2 GEMMs, K3 fused attention, K4a fused residual+dropout+LayerNorm), vocab projection + token
cross-entropy (K4b) and the integrated discriminator (K5) — runs in the sm_100a kernels behind
libsct_b200.so.  There is no CPU path: tensors must live on a B200.

Additive, keyword-only extensions (reference-preserving defaults): `fused_loss`, `return_logits`,
`compute_vuln_heads`, `greedy`, `max_new_tokens`, `n_lines`, `use_kv_cache`, `head_loss_fn`.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kernels as kn
from . import ops
from .kernels import use_ws_slot as kn_slot
from .ops import BF16, F32


class PositionalEncoding(nn.Module):
    """Sinusoidal table as a persistent buffer `pe` [max_len, 1, d] (model.py:8-21)."""

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        pos = torch.arange(max_len, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, 1, d_model)
        pe[:, 0, 0::2] = torch.sin(pos * div)
        pe[:, 0, 1::2] = torch.cos(pos * div)
        self.register_buffer("pe", pe)

    def forward(self, x):  # seq-first, as the reference calls it
        return x + self.pe[: x.size(0), :]


class ResidualLineFeatureExtractor(nn.Module):
    """model.py:128-155 (kept in PyTorch: [B, lines, d] rows only)."""

    def __init__(self, d_model):
        super().__init__()
        self.linear1 = nn.Linear(d_model, d_model)
        self.norm1 = nn.LayerNorm(d_model, eps=1e-5)
        self.linear2 = nn.Linear(d_model, d_model)
        self.norm2 = nn.LayerNorm(d_model, eps=1e-5)
        self.dropout = nn.Dropout(0.1)

    def forward(self, x):
        y = self.dropout(F.gelu(self.norm1(self.linear1(x))))
        y = self.dropout(self.norm2(self.linear2(y)))
        return y + 0.1 * x


# ---------------------------------------------------------------------------------------------------
# Sub-module shims.  inference.py does not only call `model(...)`: it drives `model.encoder(x,
# src_key_padding_mask=)`, `model.ast_attention(query=, key=, value=, key_padding_mask=) -> (out, weights)`,
# `model.cross_attention(...)` and `model.decoder(tgt, memory, tgt_mask=, memory_key_padding_mask=)` directly
# (/root/reference/SCT-GAN/inference.py:558, 563-577, 1150-1155, 1263, 1272-1277).  These subclasses keep the torch
# module classes' parameters and state_dict keys (they only override `forward`) and send those calls through the
# same sm_100a kernels as `SmartContractTransformer.forward` — no stock nn.Transformer* / cuDNN path, no CPU path.
# ---------------------------------------------------------------------------------------------------
def _owner_of(mod):
    owner = mod.__dict__.get("_sct_owner")
    owner = owner() if owner is not None else None
    if owner is None:
        raise RuntimeError("this module runs only as a sub-module of sct_gan_b200.SmartContractTransformer")
    return owner


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"sct_gan_b200 {what} runs on a B200 only (no CPU fallback): move the tensors to cuda")


class FusedTransformerEncoder(nn.TransformerEncoder):
    """`model.encoder(src, src_key_padding_mask=~mask)` (model.py:424-428, inference.py:558): [B, S, d] fp32 in/out."""

    def forward(self, src, mask=None, src_key_padding_mask=None, is_causal=None):
        owner = _owner_of(self)
        _need_cuda(src, "encoder")
        if mask is not None or is_causal:
            raise NotImplementedError("encoder attn_mask is not part of the SCT-GAN path (only src_key_padding_mask)")
        B, S, d = src.shape
        owner._begin_pass(src.device)
        kpm = src_key_padding_mask.bool().contiguous() if src_key_padding_mask is not None else None
        mem, _ = owner._encode(src.reshape(B * S, d).float().contiguous(), B, S, kpm)
        ops.DropoutRng.detach_sites(owner)
        return mem.view(B, S, d)


class FusedTransformerDecoder(nn.TransformerDecoder):
    """`model.decoder(tgt, memory, tgt_mask=causal | None, memory_key_padding_mask=~src_mask)` (model.py:949-954,
    inference.py:1150-1155, 1272-1277).  Returns the residual stream [B, T, d] fp32 WITHOUT output_norm, like
    nn.TransformerDecoder(norm=None).  tgt_mask must be None or the square-subsequent mask (float -inf / bool upper
    triangle): the kernel applies the causal predicate itself, the mask tensor is only checked."""

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, tgt_is_causal=None, memory_is_causal=False):
        owner = _owner_of(self)
        _need_cuda(tgt, "decoder")
        if memory_mask is not None or tgt_key_padding_mask is not None or memory_is_causal:
            raise NotImplementedError("decoder: only tgt_mask (causal) and memory_key_padding_mask are on the SCT-GAN path")
        B, T, d = tgt.shape
        S = memory.shape[1]
        causal = False
        if tgt_mask is not None:
            m = tgt_mask if tgt_mask.dtype == torch.bool else (tgt_mask == float("-inf"))
            want = torch.ones((T, T), dtype=torch.bool, device=m.device).triu(1)
            ok = m.shape == (T, T) and bool((m == want).all())
            if ok and tgt_mask.dtype != torch.bool:
                ok = bool((tgt_mask.masked_fill(m, 0.0) == 0).all())
            if not ok:
                raise NotImplementedError("decoder tgt_mask must be None or generate_square_subsequent_mask(T)")
            causal = True
        owner._begin_pass(tgt.device)
        _, mem_b = ops.residual_ln(memory.reshape(B * S, d).float().contiguous(), None, None, None, mode="cast", want_x=False)
        kpm = memory_key_padding_mask.bool().contiguous() if memory_key_padding_mask is not None else None
        x = owner._decode(tgt.reshape(B * T, d).float().contiguous(), mem_b, B, T, S, kpm, causal=causal,
                          final_norm=False)
        ops.DropoutRng.detach_sites(owner)
        return x.view(B, T, d)


class FusedMultiheadAttention(nn.MultiheadAttention):
    """`att(query=, key=, value=, key_padding_mask=) -> (out, None)` for ast_attention / cross_attention /
    disc_path_attention (model.py:431-448, 1186; inference.py:563-577): [B, L, d] fp32 in/out, key is value.  The
    averaged attention weights nn.MultiheadAttention would return are discarded by every caller
    (model.py:433, 443, 1186) and are never materialised here: the second element is None."""

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None,
                average_attn_weights=True, is_causal=False):
        owner = _owner_of(self)
        _need_cuda(query, "attention")
        if attn_mask is not None or is_causal:
            raise NotImplementedError("attn_mask is not part of the SCT-GAN path for this module")
        if key is not value and not (key.data_ptr() == value.data_ptr() and key.shape == value.shape):
            raise NotImplementedError("key and value must be the same tensor (as at every SCT-GAN call site)")
        B, Lq, d = query.shape
        Lk = key.shape[1]
        owner._begin_pass(query.device)
        _, qb = ops.residual_ln(query.reshape(B * Lq, d).float().contiguous(), None, None, None, mode="cast", want_x=False)
        if key is query:
            kb = qb
        else:
            _, kb = ops.residual_ln(key.reshape(B * Lk, d).float().contiguous(), None, None, None, mode="cast", want_x=False)
        kpm = key_padding_mask.bool().contiguous() if key_padding_mask is not None else None
        o = owner._mha_cross(qb, kb, self, B, Lq, Lk, kpm)
        ops.DropoutRng.detach_sites(owner)
        return o.float().view(B, Lq, d), None


class SmartContractTransformer(nn.Module):
    def __init__(self, d_model=768, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048,
                 dropout=0.3, max_length=1024, vocab_size=50265, num_vulnerability_types=8, use_gan=False):
        super().__init__()
        d = d_model
        # --- module tree: names and registration order follow model.py:40-279 (state_dict compatibility)
        self.embedding = nn.Embedding(vocab_size, d)
        self.embedding_dropout = nn.Dropout(dropout)
        self.embedding_norm = nn.LayerNorm(d)
        self.pos_encoder = PositionalEncoding(d, max_length)
        self.ast_embedding = nn.Embedding(vocab_size, d)
        self.ast_embedding_dropout = nn.Dropout(dropout)
        self.ast_embedding_norm = nn.LayerNorm(d)
        self.path_embedding = self.ast_embedding
        enc_layer = nn.TransformerEncoderLayer(d_model=d, nhead=nhead, dim_feedforward=dim_feedforward,
                                               dropout=dropout, batch_first=True, activation="gelu", norm_first=True)
        self.encoder = FusedTransformerEncoder(enc_layer, num_layers=num_encoder_layers, enable_nested_tensor=False)
        dec_layer = nn.TransformerDecoderLayer(d_model=d, nhead=nhead, dim_feedforward=dim_feedforward,
                                               dropout=dropout, batch_first=True, activation="gelu", norm_first=True)
        self.decoder = FusedTransformerDecoder(dec_layer, num_layers=num_decoder_layers)
        self.output_norm = nn.LayerNorm(d)
        self.output_dropout = nn.Dropout(dropout)
        self.output_layer = nn.Linear(d, vocab_size)
        self.contract_feature_aggregation = nn.Sequential(
            nn.Linear(2 * d, 2 * d), nn.LayerNorm(2 * d), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(2 * d, d), nn.LayerNorm(d), nn.GELU(), nn.Dropout(dropout))
        self.contract_vuln_attention = nn.MultiheadAttention(d, nhead, dropout=dropout, batch_first=True)
        self.contract_vulnerability_head = nn.Sequential(
            nn.Linear(d, d), nn.LayerNorm(d), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(d, d // 2), nn.LayerNorm(d // 2), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(d // 2, num_vulnerability_types))
        self.line_feature_extractor = ResidualLineFeatureExtractor(d)
        self.line_vuln_attention = nn.MultiheadAttention(d, nhead, dropout=dropout * 0.2, batch_first=True)
        self.vuln_type_attention = nn.MultiheadAttention(d, nhead, dropout=dropout * 0.2, batch_first=True)
        self.line_vulnerability_head_1 = nn.Sequential(
            nn.Linear(2 * d, d), nn.GELU(), nn.Dropout(0.1), nn.Linear(d, d // 2), nn.GELU(), nn.Dropout(0.1),
            nn.Linear(d // 2, num_vulnerability_types))
        self.line_specific_processor = nn.Sequential(
            nn.Linear(d, d), nn.GELU(), nn.Dropout(0.1), nn.Linear(d, d // 2), nn.GELU(), nn.Dropout(0.1))
        self.vuln_type_processor = nn.ModuleList([
            nn.Sequential(nn.Linear(d // 2, d // 4), nn.GELU(), nn.Dropout(0.1), nn.Linear(d // 4, 1))
            for _ in range(num_vulnerability_types)])
        self._debug_mode = False
        self.ast_attention = FusedMultiheadAttention(d, nhead, dropout=dropout, batch_first=True)
        self.cross_attention = FusedMultiheadAttention(d, nhead, dropout=dropout, batch_first=True)
        self.feature_fusion = nn.Sequential(
            nn.Linear(2 * d, d), nn.LayerNorm(d), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(d, d // 2), nn.LayerNorm(d // 2), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(d // 2, d))
        self.use_gan = use_gan
        if use_gan:
            self.disc_path_attention = FusedMultiheadAttention(d, nhead, dropout=dropout, batch_first=True)
            self.disc_grammar_embedding = nn.Embedding(vocab_size, d)  # never used in forward (model.py:249)
            self.disc_grammar_projection = nn.Linear(d, d)
            self.disc_feature_extractor = nn.Sequential(
                nn.Linear(d, 2 * d), nn.LayerNorm(2 * d), nn.GELU(), nn.Dropout(dropout),
                nn.Linear(2 * d, d), nn.LayerNorm(d), nn.GELU(), nn.Dropout(dropout))
            self.disc_synthetic_head = nn.Sequential(
                nn.Linear(d, d // 2), nn.LayerNorm(d // 2), nn.GELU(), nn.Dropout(dropout), nn.Linear(d // 2, 1))
        self.d_model = d
        self.nhead = nhead
        self.dropout_p = dropout
        self.max_length = max_length
        self.vocab_size = vocab_size
        self.num_vulnerability_types = num_vulnerability_types
        self.empty_line_embedding = nn.Parameter(torch.zeros(d))
        self._init_weights()
        for p in self.feature_fusion.parameters():  # model.py:285-286: clamp parameter grads to +-1
            p.register_hook(self.hook_fn)
        self._shadow = ops.ShadowCache()
        import weakref

        for sub in (self.encoder, self.decoder, self.ast_attention, self.cross_attention,
                    getattr(self, "disc_path_attention", None)):
            if sub is not None:  # plain attribute (not a registered sub-module / buffer): no cycle in state_dict
                sub.__dict__["_sct_owner"] = weakref.ref(self)
        self._heads_stream = None
        # training: vulnerability heads on a side stream, overlapped with the decoder (SCT_HEADS_SIDE_STREAM=0: A/B timing)
        self.heads_side_stream = __import__("os").environ.get("SCT_HEADS_SIDE_STREAM", "1") != "0"
        self._step_counter = 0
        self.heads_autocast = True  # bf16 matmuls for the PyTorch vulnerability heads on the GPU
        self.fused_ffn = __import__("os").environ.get("SCT_FUSED_FFN", "1") != "0"  # =0: separate gelu_dropout kernels (A/B)

    # ------------------------------------------------------------------------------------ init
    def _init_weights(self):
        """Same distributions as model.py:288-383 (every 1-D parameter zero, matrices Xavier, embeddings /
        vocab projection / contract head N(0, 0.02), line feature extractor N(0, 0.1), ...)."""
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
            else:
                nn.init.zeros_(p)
        for w in (self.embedding.weight, self.ast_embedding.weight, self.output_layer.weight):
            nn.init.normal_(w, 0.0, 0.02)
        for m in self.contract_vulnerability_head:
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0.0, 0.02)
        for lin in (self.line_feature_extractor.linear1, self.line_feature_extractor.linear2):
            nn.init.normal_(lin.weight, 0.0, 0.1)
        for att in (self.line_vuln_attention, self.vuln_type_attention):
            for p in att.parameters():
                if p.dim() > 1:
                    nn.init.xavier_uniform_(p, gain=0.8)
        last = self.line_vulnerability_head_1[-1]
        nn.init.normal_(last.weight, 0.0, 0.1)
        nn.init.constant_(last.bias, -0.2)

    def hook_fn(self, grad):
        return torch.clamp(grad, -1.0, 1.0)

    def generate_square_subsequent_mask(self, sz):
        """Float mask, 0 on/below the diagonal and -inf above (model.py:389-393)."""
        return torch.full((sz, sz), float("-inf")).triu(1)

    def set_current_epoch(self, epoch):
        self.current_epoch = epoch

    # ------------------------------------------------------------------------------ fused blocks
    def _begin_pass(self, device, whole_forward=False):
        """Start of a pass through the kernels: bf16 weight copies are re-validated; in training mode `forward`
        advances the model's dropout epoch (device counter: also advances under CUDA-graph replay) and numbers its
        dropout sites from 0.  A sub-module shim called on its own (inference.py's direct `model.encoder(...)` calls)
        must NOT move the epoch — the backward of an earlier shim call would then regenerate different masks — so
        its sites continue a separate host-side numbering instead (fresh masks per call, epoch untouched)."""
        self._shadow.begin_step(refresh=self.training)
        if not self.training:
            return
        if whole_forward:
            self._step_counter += 1
            ops.DropoutRng.begin_step(self, device)
        else:
            ops.DropoutRng.attach(self, device)

    def _p(self):
        return self.dropout_p if self.training else 0.0

    def _w(self, p):
        return self._shadow.get(p)

    def _lin(self, x, lin: nn.Linear):
        return ops.linear(x, lin.weight, lin.bias, self._w(lin.weight))

    def _lin_rows(self, x, weight, bias, r0, r1):
        """Linear with rows [r0, r1) of a packed in-projection (torch _in_projection_packed)."""
        return ops.linear(x, weight[r0:r1], bias[r0:r1], self._w(weight)[r0:r1])

    def _embed(self, ids, emb: nn.Embedding, norm: nn.LayerNorm, want_f32, want_bf16):
        B, S = ids.shape
        if S > self.pos_encoder.pe.shape[0]:
            raise RuntimeError(f"sequence length {S} exceeds max_length {self.pos_encoder.pe.shape[0]}")
        pe = self.pos_encoder.pe.view(-1, self.d_model)
        return ops.embed_ln_pe(ids, emb.weight, norm.weight, norm.bias, pe, S, math.sqrt(self.d_model), self._p(),
                               want_f32, want_bf16)

    def _mha_self(self, y, att: nn.MultiheadAttention, B, L, kpm, causal):
        qkv = ops.linear(y, att.in_proj_weight, att.in_proj_bias, self._w(att.in_proj_weight))
        o = ops.self_attention(qkv, B, self.nhead, L, kpm, causal, att.dropout if self.training else 0.0)
        return self._lin(o, att.out_proj)

    def _mha_cross(self, yq, ykv, att: nn.MultiheadAttention, B, Lq, Lk, kpm):
        q, kv = ops.cross_proj(yq, ykv, att.in_proj_weight, att.in_proj_bias, self._w(att.in_proj_weight))
        o = ops.cross_attention(q, kv, B, self.nhead, Lq, Lk, kpm, att.dropout if self.training else 0.0)
        return self._lin(o, att.out_proj)

    def _ffn(self, y, layer):
        l1, l2 = layer.linear1, layer.linear2
        if self.fused_ffn and y.shape[0] > 128 and l1.out_features % 64 == 0 and l1.out_features >= 512 \
                and l1.bias is not None and l2.bias is not None:
            # activation + dropout inside the GEMM epilogues (forward: linear1's; backward: linear2's dgrad)
            return ops.fused_ffn(y, l1, l2, self._w(l1.weight), self._w(l2.weight), self._p())
        h = ops.gelu_dropout(self._lin(y, l1), self._p())  # skinny rows (decode step) / odd widths
        return self._lin(h, l2)

    def _encode(self, x, B, S, kpm):
        """nn.TransformerEncoder, norm_first (torch transformer.py:944-983): x fp32 [B*S, d] -> memory."""
        p = self._p()
        layers = self.encoder.layers
        _, y = ops.residual_ln(x, None, layers[0].norm1.weight, layers[0].norm1.bias, mode="ln", want_x=False)
        for i, layer in enumerate(layers):
            a = self._mha_self(y, layer.self_attn, B, S, kpm, False)
            x, y = ops.residual_ln(x, a, layer.norm2.weight, layer.norm2.bias, 1.0, p, "ln")
            f = self._ffn(y, layer)
            if i + 1 < len(layers):
                nxt = layers[i + 1].norm1
                x, y = ops.residual_ln(x, f, nxt.weight, nxt.bias, 1.0, p, "ln")
            else:
                x, y = ops.residual_ln(x, f, None, None, 1.0, p, "cast")
        return x, y  # memory fp32 and its bf16 cast

    def _ast_fuse(self, mem, mem_b, ast_b, B, S, P, ast_kpm):
        """model.py:431-451: two 0.1-scaled AST attentions and the feature-fusion MLP."""
        p = self._p()
        a = self._mha_cross(mem_b, ast_b, self.ast_attention, B, S, P, ast_kpm)
        mem, mem_b = ops.residual_ln(mem, a, None, None, 0.1, 0.0, "cast")
        c = self._mha_cross(mem_b, ast_b, self.cross_attention, B, S, P, ast_kpm)
        ff = self.feature_fusion
        z = self._lin(ops.concat_scaled(mem_b, c, 0.1), ff[0])
        z = self._lin(ops.ln_act(z, ff[1].weight, ff[1].bias, p), ff[4])
        z = self._lin(ops.ln_act(z, ff[5].weight, ff[5].bias, p), ff[8])
        return ops.residual_ln(mem, z, None, None, 0.1, 0.0, "cast")

    def _decode(self, x, mem_b, B, T, S, src_kpm, causal=True, final_norm=True):
        """nn.TransformerDecoder, norm_first (torch transformer.py:1131-1205) + output_norm/output_dropout.
        Returns the bf16 rows fed to the vocab projection; with final_norm=False the fp32 residual stream of the
        last layer instead (what `model.decoder(...)` itself returns: the stack has norm=None)."""
        p = self._p()
        layers = self.decoder.layers
        _, y = ops.residual_ln(x, None, layers[0].norm1.weight, layers[0].norm1.bias, mode="ln", want_x=False)
        for i, layer in enumerate(layers):
            a = self._mha_self(y, layer.self_attn, B, T, None, causal)
            x, y = ops.residual_ln(x, a, layer.norm2.weight, layer.norm2.bias, 1.0, p, "ln")
            c = self._mha_cross(y, mem_b, layer.multihead_attn, B, T, S, src_kpm)
            x, y = ops.residual_ln(x, c, layer.norm3.weight, layer.norm3.bias, 1.0, p, "ln")
            f = self._ffn(y, layer)
            last = i + 1 == len(layers)
            if last and not final_norm:
                x, _ = ops.residual_ln(x, f, None, None, 1.0, p, "none")
                return x
            nxt = layers[i + 1].norm1 if not last else self.output_norm
            x, y = ops.residual_ln(x, f, nxt.weight, nxt.bias, 1.0, p, "ln", want_x=not last)
        if p > 0:
            _, y = ops.residual_ln(None, y, None, None, 1.0, p, "cast", want_x=False)
        return y

    def discriminator_forward(self, features, _features_bf16=None):
        """model.py:1174-1201.  `features` [B, S, d] fp32.  The grammar projection is applied after the
        sequence mean (mean and Linear commute), so it runs on [B, d] rows."""
        if not self.use_gan:
            return None
        B, S, d = features.shape
        f2 = features.reshape(B * S, d)
        fb = _features_bf16
        if fb is None:
            _, fb = ops.residual_ln(f2, None, None, None, mode="cast", want_x=False)
        a = self._mha_self(fb, self.disc_path_attention, B, S, None, False)
        pooled = ops.seq_mean(f2, a, B, S)
        p = self._p()
        x = ops.small_linear(pooled, self.disc_grammar_projection.weight, self.disc_grammar_projection.bias)
        fe, sh = self.disc_feature_extractor, self.disc_synthetic_head
        x = ops.ln_act(ops.small_linear(x, fe[0].weight, fe[0].bias, True), fe[1].weight, fe[1].bias, p)
        x = ops.ln_act(ops.small_linear(x, fe[4].weight, fe[4].bias, True), fe[5].weight, fe[5].bias, p)
        x = ops.ln_act(ops.small_linear(x, sh[0].weight, sh[0].bias, True), sh[1].weight, sh[1].bias, p)
        return ops.small_linear(x, sh[4].weight, sh[4].bias)

    # ----------------------------------------------------------------- PyTorch (out-of-scope) heads
    def _contract_heads(self, memory, mem_b=None):
        """model.py:455-476 (contract-level logits).  On the CUDA path the attention of the mean query over the
        [B, S, d] memory uses the same kernels as the hot path (K/V projection = tcgen05 GEMM on the bf16
        memory, K3 with Lq = 1); the [B, .] MLPs stay in PyTorch."""
        B, S, d = memory.shape
        att = self.contract_vuln_attention
        if mem_b is None or not memory.is_cuda:
            q = memory.mean(dim=1, keepdim=True)
            a, _ = att(query=q, key=memory, value=memory, need_weights=False)
            rep = torch.cat([memory.mean(dim=1), a.squeeze(1)], dim=-1)
        else:
            avg = ops.seq_mean(memory.reshape(B * S, d), None, B, S)
            q = ops.small_linear(avg, att.in_proj_weight[:d], att.in_proj_bias[:d], True)
            kv = self._lin_rows(mem_b, att.in_proj_weight, att.in_proj_bias, d, 3 * d)
            o = ops.cross_attention(q, kv, B, self.nhead, 1, S, None, att.dropout if self.training else 0.0)
            a = ops.small_linear(o, att.out_proj.weight, att.out_proj.bias)
            rep = torch.cat([avg, a], dim=-1)
        return self.contract_vulnerability_head(self.contract_feature_aggregation(rep))

    def _line_position_encoding(self, n_lines, device):
        pos = torch.arange(n_lines, dtype=torch.float, device=device).unsqueeze(1)
        div = torch.exp(torch.arange(0, self.d_model, 2, dtype=torch.float, device=device)
                        * -(math.log(10000.0) / self.d_model))
        pe = torch.zeros(n_lines, self.d_model, device=device)
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        return pe

    def _type_processors(self, spec):
        """The per-type heads of model.py:741-755 (`vuln_type_processor[t]` = Linear(d/2, d/4) -> GELU -> Dropout(0.1)
        -> Linear(d/4, 1), one per vulnerability type) evaluated together: the first layers are one linear with the
        weights concatenated, the second layers a weighted sum per type — ~10 kernels per pass instead of ~25 per
        type (the [B, lines, .] tensors are tiny: launch count is what they cost)."""
        procs = self.vuln_type_processor
        T = len(procs)
        w1 = torch.cat([p[0].weight for p in procs], dim=0)          # [T * d/4, d/2]
        b1 = torch.cat([p[0].bias for p in procs], dim=0)            # [T * d/4]
        w2 = torch.cat([p[3].weight for p in procs], dim=0)          # [T, d/4]
        b2 = torch.cat([p[3].bias for p in procs], dim=0)            # [T]
        h = F.dropout(F.gelu(F.linear(spec, w1, b1)), procs[0][2].p, self.training)
        h = h.view(*spec.shape[:-1], T, -1)
        return (h * w2.to(h.dtype)).sum(dim=-1) + b2.to(h.dtype)

    def _line_heads(self, memory, token_to_line, n_lines=None, mem_b=None):
        """model.py:480-759 with the two Python loops (batch x lines, lines x types) batched: every line
        goes through the same weights, so [B, L, .] tensors give the same result.  `n_lines` (=
        token_to_line.max() + 1, model.py:484) may be passed by a caller that already knows it on the host;
        otherwise it is read back from the device like the reference does."""
        B, S, d = memory.shape
        if token_to_line is not None:
            t2l = token_to_line if token_to_line.dim() == 2 else token_to_line.unsqueeze(0).expand(B, -1)
            if n_lines is None:
                n_lines = int(t2l.max().item()) + 1
            idx = torch.where((t2l >= 0) & (t2l < n_lines), t2l, torch.full_like(t2l, n_lines)).long()
            if mem_b is not None and memory.is_cuda:
                # per-line sums as a batched matmul with the 0/1 line-assignment matrix on the bf16 copy of the memory
                # (tensor cores, fp32 accumulation; the heads below run in bf16 anyway) instead of a 25 M-element atomic
                # scatter-add and its gather / index_put backward
                assign = (idx.unsqueeze(1) == torch.arange(n_lines, device=idx.device).view(1, -1, 1))  # [B, lines, S]
                cnt = assign.sum(dim=2, keepdim=True).to(memory.dtype)
                sums = torch.bmm(assign.to(mem_b.dtype), mem_b.view(B, S, d)).to(memory.dtype)
            else:
                sums = torch.zeros(B, n_lines + 1, d, device=memory.device, dtype=memory.dtype)
                sums.scatter_add_(1, idx.unsqueeze(-1).expand(-1, -1, d), memory)
                cnt = torch.zeros(B, n_lines + 1, device=memory.device, dtype=memory.dtype)
                cnt.scatter_add_(1, idx, torch.ones_like(idx, dtype=memory.dtype))
                sums, cnt = sums[:, :n_lines], cnt[:, :n_lines].unsqueeze(-1)
            mean = sums / cnt.clamp(min=1.0)
            line_features = torch.where(cnt > 0, mean, self.empty_line_embedding.expand(B, n_lines, d))
            line_features = line_features + self._line_position_encoding(n_lines, memory.device)
        else:
            line_features = memory
        original = line_features
        # On the GPU these [B, lines, .] PyTorch heads run their matmuls in bf16 (autocast), like the rest of the
        # step; LayerNorm / softmax statistics stay fp32.  On CPU (tests against the oracle) they are plain fp32.
        with torch.autocast("cuda", dtype=BF16, enabled=memory.is_cuda and self.heads_autocast):
            lf = self.line_feature_extractor(line_features)
            lf = torch.where(lf.float().std() < 1e-6, (original * 0.1).to(lf.dtype), lf)
            att1, _ = self.line_vuln_attention(lf, lf, lf, need_weights=False)
            lf = lf + 0.05 * att1
            att2, _ = self.vuln_type_attention(lf, lf, lf, need_weights=False)
            lf = lf + 0.05 * att2
            main = self.line_vulnerability_head_1(torch.cat([lf, att1], dim=-1))
            spec = self.line_specific_processor(original)
            typed = self._type_processors(spec)
            logits = (main + 0.1 * typed).float()
        n = logits.shape[1]
        if n < 1024:
            logits = torch.cat([logits, logits.new_zeros(B, 1024 - n, logits.shape[2])], dim=1)
        elif n > 1024:
            logits = logits[:, :1024]
        return logits

    # --------------------------------------------------------------------------------- forward
    def forward(self, input_ids, attention_mask=None, ast_input_ids=None, ast_attention_mask=None,
                target_ids=None, token_to_line=None, apply_syntax_constraints=True, *, fused_loss=False,
                return_logits=True, compute_vuln_heads=True, greedy=False, max_new_tokens=None, n_lines=None,
                use_kv_cache=True, head_loss_fn=None):
        if not input_ids.is_cuda:
            raise RuntimeError("sct_gan_b200 runs on a B200 only (no CPU fallback): move the batch to cuda")
        B, S = input_ids.shape
        d = self.d_model
        self._begin_pass(input_ids.device, whole_forward=True)
        x, _ = self._embed(input_ids, self.embedding, self.embedding_norm, True, False)
        src_mask = attention_mask.bool() if attention_mask is not None else \
            torch.ones((B, S), dtype=torch.bool, device=input_ids.device)
        src_kpm = (~src_mask).contiguous()
        mem, mem_b = self._encode(x, B, S, src_kpm)
        if ast_attention_mask is not None:
            P = ast_input_ids.shape[1]
            _, ast_b = self._embed(ast_input_ids, self.ast_embedding, self.ast_embedding_norm, False, True)
            ast_kpm = (~ast_attention_mask.bool()).contiguous()
            mem, mem_b = self._ast_fuse(mem, mem_b, ast_b, B, S, P, ast_kpm)
        memory = mem.view(B, S, d)

        heads_stream = None
        if compute_vuln_heads and self.training and self.heads_side_stream and target_ids is not None:
            # The vulnerability heads are ~450 tiny kernels on [B, .] / [B, lines, .] rows (3 ms of launch latency, a few
            # SMs): run them on a side stream next to the decoder.  Autograd runs a node's backward on the stream of its
            # forward, so their backward overlaps with the decoder's as well (inside the step's CUDA graph the two
            # branches are independent until the encoder needs d(memory)).  Joined before the dict is returned.
            main = torch.cuda.current_stream()
            if self._heads_stream is None or self._heads_stream.device != memory.device:
                self._heads_stream = torch.cuda.Stream(device=memory.device)
            heads_stream = self._heads_stream
            heads_stream.wait_stream(main)
            with torch.cuda.stream(heads_stream), kn_slot(1):
                contract_logits = self._contract_heads(memory, mem_b)
                line_logits = self._line_heads(memory, token_to_line, n_lines, mem_b)
                # the trainer's vulnerability-head losses (dozens of tiny element-wise kernels, forward and backward)
                # belong to this branch too: computed here they overlap with the decoder in both directions instead of
                # sitting between the end of forward and the start of backward on the main stream
                head_losses = head_loss_fn(contract_logits, line_logits) if head_loss_fn is not None else None
            mem.record_stream(heads_stream)
            mem_b.record_stream(heads_stream)
        elif compute_vuln_heads:
            contract_logits = self._contract_heads(memory, mem_b)
            line_logits = self._line_heads(memory, token_to_line, n_lines, mem_b)
            head_losses = head_loss_fn(contract_logits, line_logits) if head_loss_fn is not None else None
        else:
            contract_logits = line_logits = head_losses = None

        if target_ids is None:
            seq = self._generate(mem_b, B, S, src_kpm, apply_syntax_constraints, greedy, max_new_tokens, use_kv_cache)
            return {"generated_sequence": seq, "contract_vulnerability_logits": contract_logits,
                    "line_vulnerability_logits": line_logits}

        T = target_ids.shape[1]
        tx, _ = self._embed(target_ids, self.embedding, self.embedding_norm, True, False)
        h = self._decode(tx, mem_b, B, T, S, src_kpm)
        shifted = target_ids[:, 1:].contiguous().view(-1)
        out = {"target_ids": shifted}
        ol = self.output_layer
        if fused_loss:
            tgt = torch.full((B, T), -1, dtype=torch.long, device=target_ids.device)
            tgt[:, :-1] = target_ids[:, 1:]
            loss, lse = ops.vocab_ce(h, ol.weight, ol.bias, self._w(ol.weight), tgt.view(-1), B * (T - 1))
            out["gen_ce_loss"] = loss
            out["lse"] = lse.view(B, T)[:, :-1].reshape(-1)
        if return_logits and not fused_loss:
            logits = ops.linear(h, ol.weight, ol.bias, self._w(ol.weight))
            out["logits"] = logits.view(B, T, -1)[:, :-1, :].float().reshape(B * (T - 1), -1)
        elif return_logits:
            with torch.no_grad():
                logits = ops.linear(h.detach(), ol.weight, ol.bias, self._w(ol.weight))
            out["logits"] = logits.view(B, T, -1)[:, :-1, :].float().reshape(B * (T - 1), -1)
        if heads_stream is not None:  # join: whoever consumes the logits does so on the caller's stream
            main = torch.cuda.current_stream()
            main.wait_stream(heads_stream)
            contract_logits.record_stream(main)
            line_logits.record_stream(main)
        out["contract_vulnerability_logits"] = contract_logits
        out["line_vulnerability_logits"] = line_logits
        if head_losses is not None:
            out["head_losses"] = head_losses
        out["encoder_output"] = memory.mean(dim=1)
        out["discriminator_logits"] = self.discriminator_forward(memory, mem_b) if self.use_gan else None
        return out

    # ------------------------------------------------------------------------------ generation
    def _apply_syntax_constraints(self, logits, prev_tokens):
        """Only live effect of model.py:975-1060: double the ';' logit (id 59) after ids 2000-2002."""
        last = prev_tokens[:, -1]
        hit = (last >= 2000) & (last <= 2002)
        if logits.size(1) > 59:
            logits = logits.clone()
            logits[:, 59] = torch.where(hit, logits[:, 59] * 2.0, logits[:, 59])
        return logits

    def _sample_rng(self, dev):
        """Draw counter of the sampling kernel on `dev` (a device int64 the kernel reads at run time, so a captured
        decode step draws new numbers on every replay).  It restarts from a hash of torch's seed whenever
        `torch.manual_seed` has been called since the last look — generation is reproducible under a fixed seed."""
        state = self.__dict__.setdefault("_sample_state", {})
        st = state.get(dev)
        if st is None:
            st = state[dev] = {"ctr": torch.zeros(1, dtype=torch.long, device=dev), "seed": None}
        if not torch.cuda.is_current_stream_capturing() and st["seed"] != torch.initial_seed():
            st["seed"] = torch.initial_seed()
            st["ctr"].fill_((st["seed"] * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019) & ((1 << 62) - 1))
        return st

    def _sample(self, logits, tgt, apply_syntax_constraints, greedy):
        """model.py:892-918: temperature 0.7, syntax tweak, top-k 50, top-p 0.95, multinomial (or argmax) — one kernel
        over the bf16 logits rows (csrc/sample.cu, `sct_sample_rows`).  Returns int64 [rows, 1]."""
        V = logits.shape[-1]
        if logits.dtype != BF16 or logits.stride(-1) != 1 or logits.stride(0) % 8 or logits.data_ptr() % 16:
            buf = torch.empty((logits.shape[0], (V + 7) // 8 * 8), dtype=BF16, device=logits.device)[:, :V]
            buf.copy_(logits)
            logits = buf
        prev = tgt[:, -1].contiguous() if apply_syntax_constraints else None
        if greedy:
            return kn.sample_rows(logits, V, prev, greedy=True)
        st = self._sample_rng(logits.device)
        nxt = kn.sample_rows(logits, V, prev, temperature=0.7, top_k=50, top_p=0.95, greedy=False,
                             seed=st["seed"] & ((1 << 63) - 1), offset=0, epoch=st["ctr"])
        st["ctr"].add_(1)
        return nxt

    @staticmethod
    def _stop_update(stop_at, nxt, pos):
        """The same rules as `_stop` evaluated on the device for step i = pos: (EOS or PAD anywhere and i > 50) or (every
        sequence emitted EOS and i > 20); `stop_at` keeps the first step at which they held."""
        eos = nxt == 2
        cond = ((eos.any() | (nxt == 0).any()) & (pos > 50)) | (eos.all() & (pos > 20))
        stop_at.copy_(torch.minimum(stop_at, torch.where(cond, pos, stop_at)))

    def _stop(self, nxt, i):
        """model.py:923-930 (batch-wide early stop; a host decision in the reference as well)."""
        stop = ((nxt == 2).any() | (nxt == 0).any()).item()
        return (stop and i > 50) or (i > 20 and bool((nxt == 2).all().item()))

    @torch.no_grad()
    def _generate(self, mem_b, B, S, src_kpm, apply_syntax_constraints, greedy, max_new_tokens, use_kv_cache=True):
        """The sampling loop of model.py:862-930 (BOS = 1) with a KV cache: the reference re-embeds and re-decodes the
        whole prefix for every new token (O(T^2) decoder passes); here each step runs the decoder on ONE position,
        attending to cached self-attention K/V ([B, T_max, 2d] per layer) and to cross-attention K/V of the encoder
        memory projected once.  Same arithmetic per position, so greedy tokens match the reference.

        The step has no host-visible state: the position is a device scalar (token / positional-encoding row / cache
        slot are selected with it, keys past it are masked by a [B, T_max] byte mask that the attention kernel turns
        into a tile count on the device), so ONE captured CUDA graph serves every step of every call with the same
        (B, S, T_max) signature."""
        if not use_kv_cache:
            return self._generate_recompute(mem_b, B, S, src_kpm, apply_syntax_constraints, greedy, max_new_tokens)
        dev, d = mem_b.device, self.d_model
        max_len = min(self.max_length, 1024)
        steps = max_len - 1 if max_new_tokens is None else min(max_len - 1, max_new_tokens)
        t_max = (steps + 127) // 128 * 128
        early_stop = max_new_tokens is None  # the reference's batch-wide stop rules (model.py:923-930)
        key = (B, S, t_max, bool(apply_syntax_constraints), bool(greedy), dev.index, early_stop)
        cache = self.__dict__.setdefault("_decode_cache", {})
        ent = cache.get(key)
        layers = self.decoder.layers
        if ent is None:
            st = {
                "pos": torch.zeros(1, dtype=torch.long, device=dev),
                "tgt": torch.ones((B, t_max + 1), dtype=torch.long, device=dev),
                "kpm_self": torch.ones((B, t_max), dtype=torch.uint8, device=dev),
                "self_kv": [torch.zeros((B, t_max, 2 * d), dtype=BF16, device=dev) for _ in layers],
                "mem_kv": [torch.empty((B * S, 2 * d), dtype=BF16, device=dev) for _ in layers],
                "src_kpm": torch.empty((B, S), dtype=torch.bool, device=dev),
                "nxt": torch.zeros((B, 1), dtype=torch.long, device=dev),
                # first step at which the stop rules held (t_max + 1 = never): evaluated on the device every step, read
                # back every 16 steps instead of two .item() round trips per token
                "stop_at": torch.full((1,), t_max + 1, dtype=torch.long, device=dev) if early_stop else None,
            }
            ent = cache[key] = {"st": st, "graph": None, "calls": 0}
        st = ent["st"]
        # per-call state: BOS, nothing cached yet, this call's encoder memory and padding mask
        st["pos"].zero_()
        st["tgt"].fill_(1)
        st["kpm_self"].fill_(1)
        st["src_kpm"].copy_(src_kpm)
        if early_stop:
            st["stop_at"].fill_(t_max + 1)
        self._shadow.begin_step(refresh=False)
        if not greedy:
            self._sample_rng(dev)  # a torch.manual_seed since the last call restarts the draw counter (outside the graph)
        for li, l in enumerate(layers):
            ca = l.multihead_attn
            st["mem_kv"][li].copy_(self._lin_rows(mem_b, ca.in_proj_weight, ca.in_proj_bias, d, 3 * d))
        for l in layers:  # make every bf16 weight shadow the step reads valid before (re)playing the graph
            for w in (l.self_attn.in_proj_weight, l.self_attn.out_proj.weight, l.multihead_attn.in_proj_weight,
                      l.multihead_attn.out_proj.weight, l.linear1.weight, l.linear2.weight):
                self._w(w)
        self._w(self.output_layer.weight)
        n_out = steps + 1
        for i in range(steps):
            if ent["graph"] is None and ent["calls"] >= 1 and i == 0:
                # second call with this signature: capture the step once (the first call was the eager warm-up)
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    self._decode_step(st, B, S, apply_syntax_constraints, greedy)
                ent["graph"] = g
            if ent["graph"] is not None:
                ent["graph"].replay()
            else:
                self._decode_step(st, B, S, apply_syntax_constraints, greedy)
            if early_stop and (i % 16 == 15 or i == steps - 1):
                hit = int(st["stop_at"].item())  # tokens decoded past the stop step are simply not returned
                if hit <= i:
                    n_out = hit + 2
                    break
        ent["calls"] += 1
        return st["tgt"][:, :n_out].clone()

    def _decode_step(self, st, B, S, apply_syntax_constraints, greedy):
        """One position through the decoder against the caches; everything indexed by the device scalar st['pos']."""
        d, H = self.d_model, self.nhead
        layers = self.decoder.layers
        pos = st["pos"]
        ids = st["tgt"].index_select(1, pos)                               # [B, 1] current input token
        pe_row = self.pos_encoder.pe.view(-1, d).index_select(0, pos)      # [1, d]
        st["kpm_self"].index_fill_(1, pos, 0)                              # the new position becomes visible
        x, _ = ops.embed_ln_pe(ids, self.embedding.weight, self.embedding_norm.weight, self.embedding_norm.bias,
                               pe_row, 1, math.sqrt(d), 0.0, True, False)
        _, y = ops.residual_ln(x, None, layers[0].norm1.weight, layers[0].norm1.bias, mode="ln", want_x=False)
        for li, layer in enumerate(layers):
            sa = layer.self_attn
            qkv = ops.linear(y, sa.in_proj_weight, sa.in_proj_bias, self._w(sa.in_proj_weight))
            st["self_kv"][li].index_copy_(1, pos, qkv[:, d:].unsqueeze(1))
            a = ops.cached_attention(qkv[:, :d], st["self_kv"][li], B, H, st["self_kv"][li].shape[1], st["kpm_self"])
            x, y = ops.residual_ln(x, self._lin(a, sa.out_proj), layer.norm2.weight, layer.norm2.bias, 1.0, 0.0, "ln")
            ca = layer.multihead_attn
            q = self._lin_rows(y, ca.in_proj_weight, ca.in_proj_bias, 0, d)
            c = ops.cross_attention(q, st["mem_kv"][li], B, H, 1, S, st["src_kpm"], 0.0)
            x, y = ops.residual_ln(x, self._lin(c, ca.out_proj), layer.norm3.weight, layer.norm3.bias, 1.0, 0.0, "ln")
            f = self._ffn(y, layer)
            nxt_norm = layers[li + 1].norm1 if li + 1 < len(layers) else self.output_norm
            x, y = ops.residual_ln(x, f, nxt_norm.weight, nxt_norm.bias, 1.0, 0.0, "ln")
        ol = self.output_layer
        logits = ops.linear(y, ol.weight, ol.bias, self._w(ol.weight))
        nxt = self._sample(logits, ids, apply_syntax_constraints, greedy)
        st["nxt"].copy_(nxt)
        st["tgt"].index_copy_(1, pos + 1, nxt)
        if st.get("stop_at") is not None:
            self._stop_update(st["stop_at"], nxt, pos)
        pos.add_(1)

    @torch.no_grad()
    def _generate_recompute(self, mem_b, B, S, src_kpm, apply_syntax_constraints, greedy, max_new_tokens):
        """The reference's own schedule (whole prefix re-decoded per token) on the fused decoder; kept as the
        cross-check of the KV-cached path."""
        dev = mem_b.device
        tgt = torch.ones((B, 1), dtype=torch.long, device=dev)
        max_len = min(self.max_length, 1024)
        steps = max_len - 1 if max_new_tokens is None else min(max_len - 1, max_new_tokens)
        ol = self.output_layer
        for i in range(steps):
            T = tgt.shape[1]
            tx, _ = self._embed(tgt, self.embedding, self.embedding_norm, True, False)
            h = self._decode(tx, mem_b, B, T, S, src_kpm)
            last = h.view(B, T, -1)[:, -1, :].contiguous()
            logits = ops.linear(last, ol.weight, ol.bias, self._w(ol.weight))
            nxt = self._sample(logits, tgt, apply_syntax_constraints, greedy)
            tgt = torch.cat([tgt, nxt], dim=1)
            if max_new_tokens is None and self._stop(nxt, i):
                break
        return tgt
