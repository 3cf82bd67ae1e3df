"""CPU suite: the oracle against the committed golden fixtures (reference outputs, oracle/make_golden.py),
the drop-in boundary (state_dict keys/shapes, constructor, C-ABI symbols) and host logic."""
import ctypes
import glob
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.pt")))


def load_golden(path):
    return torch.load(path, map_location="cpu", weights_only=False)


def test_golden_present():
    assert len(GOLDEN) >= 2


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_reference_outputs(path):
    """fp32 oracle vs fp32 reference outputs: same arithmetic, different summation order -> 2e-4 relative."""
    from oracle import sct_oracle as O

    g = load_golden(path)
    sd = O.synth_state_dict(g["shapes"], g["seed"])
    batch = O.make_batch(**g["batch_args"])
    out = O.forward_train(sd, g["cfg"], batch, torch.float32)
    assert torch.equal(out["target_ids"], g["outputs"]["target_ids"])  # integer work: bit-exact
    for k in ("logits", "contract_vulnerability_logits", "encoder_output", "discriminator_logits"):
        ref = g["outputs"][k]
        assert (out[k] - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item()), k
    n_lines = g["outputs"]["line_vulnerability_logits"].shape[1]
    lv = out["line_vulnerability_logits"]
    assert (lv[:, :n_lines] - g["outputs"]["line_vulnerability_logits"]).abs().max().item() <= 2e-4
    assert lv[:, n_lines:].abs().max().item() == 0.0  # padded to 1024 with zeros (model.py:751-757)
    losses = O.step_losses(out, batch, g["hp"])
    for k in ("gen_ce_loss", "contract_vuln_loss", "line_vuln_loss", "discriminator_loss", "total_loss"):
        assert abs(float(losses[k]) - g["losses"][k]) <= 2e-4 * max(1.0, abs(g["losses"][k])), k


@pytest.mark.parametrize("path", GOLDEN[:1], ids=[os.path.basename(p) for p in GOLDEN[:1]])
def test_oracle_gradients_match_reference(path):
    from oracle import sct_oracle as O

    g = load_golden(path)
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and k not in ("pos_encoder.pe",))
          for k, v in O.synth_state_dict(g["shapes"], g["seed"]).items()}
    sd["path_embedding.weight"] = sd["ast_embedding.weight"]
    batch = O.make_batch(**g["batch_args"])
    out = O.forward_train(sd, g["cfg"], batch, torch.float32)
    O.step_losses(out, batch, g["hp"])["total_loss"].backward()
    for n, ref in g["grads"].items():
        assert (sd[n].grad - ref).norm().item() <= 1e-3 * ref.norm().item() + 1e-9, n
    for n, ref in g["grad_rows"].items():
        assert (sd[n].grad[:4] - ref).norm().item() <= 1e-3 * ref.norm().item() + 1e-9, n
    # parameters the reference leaves without gradient
    assert g["grad_norms"]["disc_grammar_embedding.weight"] is None
    assert sd["disc_grammar_embedding.weight"].grad is None


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_greedy_generation(path):
    from oracle import sct_oracle as O

    g = load_golden(path)
    sd = O.synth_state_dict(g["shapes"], g["seed"])
    batch = O.make_batch(**g["batch_args"])
    toks, _ = O.generate_greedy(sd, g["cfg"], batch, g["greedy_tokens"].shape[1] - 1)
    assert torch.equal(toks, g["greedy_tokens"])  # token ids: bit-exact


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_state_dict_is_drop_in(path):
    """Same keys and shapes as the reference's state_dict (recorded in the fixture), strict load works."""
    from oracle import sct_oracle as O
    from sct_gan_b200 import SmartContractTransformer

    g = load_golden(path)
    m = SmartContractTransformer(**g["cfg"])
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert mine == g["shapes"]
    m.load_state_dict(O.synth_state_dict(g["shapes"], g["seed"]), strict=True)
    assert m.path_embedding is m.ast_embedding
    assert [n for n, _ in m.named_buffers()] == ["pos_encoder.pe"]


def test_default_constructor_matches_reference_inventory():
    """312 state_dict keys / 262,601,074 parameters with use_gan=True; every 1-D parameter is
    zero-initialised except the line head's output bias (model.py:290-294, 369)."""
    from sct_gan_b200 import SmartContractTransformer

    m = SmartContractTransformer(use_gan=True)
    assert len(m.state_dict()) == 312
    assert sum(p.numel() for p in m.parameters()) == 262_601_074
    for n, p in m.named_parameters():
        if p.dim() == 1 and n != "line_vulnerability_head_1.6.bias":
            assert float(p.abs().max()) == 0.0, n
    assert torch.allclose(m.line_vulnerability_head_1[6].bias, torch.full((8,), -0.2))
    for p in m.feature_fusion.parameters():  # the +-1 clamp hook of model.py:285-286
        assert p._backward_hooks is not None and len(p._backward_hooks) == 1
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 8, dtype=torch.long))  # no CPU fallback


def test_param_groups_match_reference_rules():
    """train.py:518-527 -> 245 / 22 / 22 / 21 tensors."""
    from oracle import sct_oracle as O
    from sct_gan_b200 import SmartContractTransformer
    from sct_gan_b200.trainer import param_group_of

    m = SmartContractTransformer(num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=64, vocab_size=64,
                                 max_length=16, use_gan=True)
    names = [n for n, _ in m.named_parameters()]
    mine = [param_group_of(n, True) for n in names]
    theirs = [O.param_groups(names)[n][0] for n in names]
    assert mine == theirs
    full = SmartContractTransformer(use_gan=True)
    cnt = [0, 0, 0, 0]
    for n, _ in full.named_parameters():
        cnt[param_group_of(n, True)] += 1
    assert cnt == [245, 22, 22, 21]


def test_vectorised_spatial_penalty_matches_loop():
    """trainer.spatial_penalty (O(N)) against the oracle's restated B*1024-iteration loop (train.py:174-245)."""
    from oracle import sct_oracle as O
    from sct_gan_b200.trainer import spatial_aware_focal_loss, spatial_penalty

    g = torch.Generator().manual_seed(5)
    n, C = 2 * 1024, 8
    pred = torch.randn(n, C, generator=g, dtype=torch.float64)
    target = (torch.rand(n, C, generator=g) < 0.02).double()
    t2l = (torch.arange(1024) // 7).repeat(2)
    a = spatial_penalty(pred, target, t2l)
    b = O.spatial_penalty(pred, target, t2l)
    assert (a - b).abs().max().item() < 1e-12
    assert abs(float(spatial_aware_focal_loss(pred, target, t2l, 0.25, 2.0, 0.2))
               - float(O.spatial_focal_loss(pred, target, t2l, 0.25, 2.0, 0.2))) < 1e-12
    # S != 1024: the reference silently returns zeros (train.py:187-216)
    assert spatial_penalty(pred, target, t2l[:100]).abs().max().item() == 0.0
    assert O.spatial_penalty(pred, target, t2l[:100]).abs().max().item() == 0.0


def test_cabi_library_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/sct_b200.h declares."""
    from sct_gan_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "sct_b200.h")).read()
    declared = set(re.findall(r"\b(sct_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.exported_symbols())
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().sct_version() == 200


def test_line_heads_vectorised_match_oracle_loops():
    """model._line_heads (batched PyTorch, runs on CPU) vs the oracle's restated loops (model.py:480-759)."""
    from oracle import sct_oracle as O
    from sct_gan_b200 import SmartContractTransformer

    g = load_golden(GOLDEN[0])
    cfg = g["cfg"]
    m = SmartContractTransformer(**cfg).eval()
    sd = O.synth_state_dict(g["shapes"], g["seed"])
    m.load_state_dict(sd)
    gen = torch.Generator().manual_seed(3)
    memory = torch.randn(2, 40, cfg["d_model"], generator=gen)
    t2l = torch.tensor([[0] * 5 + [1] * 10 + [3] * 20 + [4] * 5, [0] * 20 + [2] * 10 + [6] * 10])  # empty lines 2/5, 1/3-5
    with torch.no_grad():
        mine = m._line_heads(memory, t2l)
        ref = O.line_heads(sd, cfg, memory, t2l, torch.float32)
        assert (mine - ref).abs().max().item() < 1e-4
        mine_c = m._contract_heads(memory)
        ref_c = O.contract_heads(sd, cfg, memory, torch.float32)
        assert (mine_c - ref_c).abs().max().item() < 1e-4
        assert (m._line_heads(memory, None) - O.line_heads(sd, cfg, memory, None, torch.float32)).abs().max().item() < 1e-4


def test_type_processors_batched_equal_per_module_loop():
    """model._type_processors (one concatenated linear + weighted sums) == the reference's per-type Sequential loop
    (model.py:741-755), in eval mode and fp32."""
    from sct_gan_b200 import SmartContractTransformer

    torch.manual_seed(5)
    m = SmartContractTransformer(d_model=96, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=64,
                                 max_length=32, vocab_size=50).eval()
    with torch.no_grad():
        for p in m.vuln_type_processor.parameters():
            p.normal_(0.0, 0.3)
        spec = torch.randn(3, 7, 48)
        loop = torch.cat([proc(spec) for proc in m.vuln_type_processor], dim=-1)
        assert (m._type_processors(spec) - loop).abs().max().item() < 1e-5


def test_vocab_chunk_rows_fills_dgrad_waves(monkeypatch):
    """ops._vocab_chunk_rows: fewest dgrad waves for the 74 CTA pairs of a B200, scratch capped at 1 GiB."""
    from sct_gan_b200 import ops

    class Props:
        multi_processor_count = 148

    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda dev: Props())
    assert ops._vocab_chunk_rows(32768, 768, 50265, "cuda") == 5464    # 6 chunks x 66 of 74 pairs
    assert ops._vocab_chunk_rows(4096, 768, 50265, "cuda") == 4096     # one chunk, one wave
    assert ops._vocab_chunk_rows(200, 768, 777, "cuda") == 200
    big = ops._vocab_chunk_rows(1 << 20, 768, 50265, "cuda")
    assert big * 50272 * 2 <= 1 << 30 and big % 8 == 0


def test_cabi_argument_errors_are_reported_without_a_gpu():
    """Error behaviour of the boundary: bad arguments return non-zero and leave a message in sct_last_error()
    (checked before any CUDA call, so this runs on a CPU box); the Python layer turns them into RuntimeError."""
    import ctypes as C

    from sct_gan_b200 import _lib

    lib = _lib.load()
    rc = lib.sct_embed_ln_pe_fwd(None, None, None, None, None, None, None, None, 8, 4, 10, 768, 1.0, 0.0, 0, 0, None, None)
    assert rc != 0 and "null pointer" in _lib.last_error()
    buf = (C.c_float * 8)()
    p = C.cast(buf, C.c_void_p)
    rc = lib.sct_gemm_bf16_nt(p, 8, p, 8, p, 8, None, 1.0, 0, 128, 64, 128, None)
    assert rc != 0 and "empty GEMM" in _lib.last_error()
    rc = lib.sct_attn_fwd(p, 768, p, p, 768, p, 768, None, None, 1, 8, 16, 16, 64, 0, 0.125, 0.0, 0, 0, None, None)
    assert rc != 0 and "head_dim" in _lib.last_error()
    rc = lib.sct_add_dropout_ln_fwd(p, None, 1.0, None, None, p, None, None, None, 4, 100, 0.0, 0, 0, None, None)
    assert rc != 0 and "unsupported row width" in _lib.last_error()
    rc = lib.sct_ce_rows(p, p, p, p, 4, 10, 9, 1.0, 0, None)
    assert rc != 0 and "pitch" in _lib.last_error()
    with pytest.raises(RuntimeError, match="sct_colsum_bf16 failed"):
        _lib.call("sct_colsum_bf16", p, 7, p, 4, 8, 1.0, None)
    # attention backward with a caller-owned workspace: the size contract is checked before anything is launched
    need = lib.sct_attn_bwd_workspace_bytes(2, 8, 100, 300)
    assert need == 2 * 8 * 300 * 128 * 2  # [B*H][Lk][Lq rounded up to 64] bf16
    rc = lib.sct_attn_bwd_ws(p, 768, p, p, 768, p, p, 768, p, p, p, 768, p, p, 768, None, 2, 8, 100, 300, 96, 0, 0.1,
                             0.0, 0, 0, None, p, need - 1, None)
    assert rc != 0 and "workspace too small" in _lib.last_error()


def test_eval_after_training_recasts_weight_shadows():
    """ShadowCache rules (host logic): copies made during a training pass are never reused by an eval pass (fused
    optimisers update weights without bumping Tensor._version); eval copies are reused until the version moves."""
    from sct_gan_b200 import ops

    calls = []
    real = ops.kn.cast_scale
    ops.kn.cast_scale = lambda src, dst, col_off=0, scale=1.0: calls.append(1) or dst.copy_(src)
    try:
        w = torch.nn.Parameter(torch.randn(4, 8))
        sc = ops.ShadowCache()
        sc.begin_step(refresh=True)
        a = sc.get(w)
        sc.get(w)
        assert len(calls) == 1  # once per training pass
        sc.begin_step(refresh=True)
        sc.get(w)
        assert len(calls) == 2  # and again in the next one
        with torch.no_grad():
            w.data.add_(1.0)  # what a fused optimiser does: no version bump
        sc.begin_step(refresh=False)
        b = sc.get(w)
        assert len(calls) == 3 and b.data_ptr() == a.data_ptr()  # eval after training: re-cast, same buffer
        assert torch.equal(b.float(), w.detach().to(torch.bfloat16).float())
        sc.begin_step(refresh=False)
        sc.get(w)
        assert len(calls) == 3  # eval copy reused
        with torch.no_grad():
            w.mul_(2.0)  # versioned in-place update (load_state_dict, manual edits)
        sc.get(w)
        assert len(calls) == 4
    finally:
        ops.kn.cast_scale = real


def test_weight_shadows_kept_in_step_by_the_optimiser_are_not_recast():
    """ShadowCache + FusedClipAdamW contract (host logic): a copy the optimiser kernel rewrote (`mark_synced`) is used
    as is by the next training and eval passes; a versioned update of the parameter invalidates it again."""
    from sct_gan_b200 import ops

    calls = []
    real = ops.kn.cast_scale
    ops.kn.cast_scale = lambda src, dst, col_off=0, scale=1.0: calls.append(1) or dst.copy_(src)
    try:
        w = torch.nn.Parameter(torch.randn(4, 8))
        other = torch.nn.Parameter(torch.randn(2, 2))
        sc = ops.ShadowCache()
        assert sc.peek_ptr(w) == 0  # no copy yet: the optimiser table gets a null shadow pointer
        sc.begin_step(refresh=True)
        a = sc.get(w)
        assert len(calls) == 1 and sc.peek_ptr(w) == a.data_ptr()
        with torch.no_grad():  # what csrc/optim.cu does: weight and copy rewritten together, no version bump
            w.data.add_(1.0)
            a.copy_(w.detach())
        sc.mark_synced([w, other])  # parameters without a copy are ignored
        sc.begin_step(refresh=True)
        assert sc.get(w).data_ptr() == a.data_ptr() and len(calls) == 1  # next training pass: no cast
        sc.begin_step(refresh=False)
        assert sc.get(w).data_ptr() == a.data_ptr() and len(calls) == 1  # eval: still current
        with torch.no_grad():
            w.mul_(2.0)  # versioned update from outside (load_state_dict, manual edit)
        sc.begin_step(refresh=True)
        sc.get(w)
        assert len(calls) == 2  # stale: re-cast, and no longer marked as synced
        sc.begin_step(refresh=True)
        sc.get(w)
        assert len(calls) == 3
    finally:
        ops.kn.cast_scale = real


def test_line_metrics_device_predicates_match_reference_cascade():
    """trainer.line_vulnerability_metrics (no host decision) against the oracle's literal restatement of the
    adaptive-threshold cascade of train.py:1043-1140, on inputs that reach every branch."""
    from oracle import sct_oracle as O
    from sct_gan_b200.trainer import line_vulnerability_metrics

    g = torch.Generator().manual_seed(11)
    B, L, T = 4, 1024, 8
    cases = {
        "ordinary": torch.randn(B, L, T, generator=g),
        "negative logits": torch.randn(B, L, T, generator=g) - 3.0,
        "too many (conservative, then ultra)": torch.randn(B, L, T, generator=g) * 0.05 + 2.0,
        "nothing above the 0.3 floor, fallback": torch.full((B, L, T), -0.9) + 0.01 * torch.randn(B, L, T, generator=g),
        "tiny probabilities, ultra fallback": torch.full((B, L, T), -6.0) + 0.01 * torch.randn(B, L, T, generator=g),
    }
    for name, logits in cases.items():
        for transposed in (False, True):
            vl = (torch.rand(B, L, T, generator=g) < 0.02).float()
            vl_in = vl.transpose(1, 2).contiguous() if transposed else vl  # the data set ships [B, types, lines]
            ref = O.line_metrics_loops(logits, vl_in)
            got = line_vulnerability_metrics(logits, vl_in)
            mine = (got["line_vuln_accuracy"].item(), got["line_vuln_precision"].item(), got["line_vuln_recall"].item(),
                    got["line_vuln_threshold"].item(), float(got["line_vuln_predictions"].item()))
            for a_, b_ in zip(mine, ref):
                assert abs(a_ - b_) < 1e-6 * max(1.0, abs(b_)), (name, transposed, mine, ref)


def test_device_stop_rule_matches_host_rule():
    """model._stop_update (device-side early stop of the KV-cached generation loop) finds the same first stop step as
    the reference's per-token host rule `_stop` (model.py:923-930)."""
    from sct_gan_b200 import SmartContractTransformer

    m = SmartContractTransformer.__new__(SmartContractTransformer)  # only the two rule methods are used
    g = torch.Generator().manual_seed(2)
    for trial in range(40):
        B, steps = 3, 90
        toks = torch.randint(3, 50, (steps, B, 1), generator=g)
        kind = trial % 4
        if kind == 0:    # one sequence emits EOS late: stops at the first such step after 50
            toks[int(torch.randint(40, 80, (1,), generator=g)), 1, 0] = 2
        elif kind == 1:  # everybody emits EOS between 21 and 50
            toks[int(torch.randint(15, 45, (1,), generator=g)), :, 0] = 2
        elif kind == 2:  # a PAD token
            toks[int(torch.randint(30, 85, (1,), generator=g)), 2, 0] = 0
        host = next((i for i in range(steps) if m._stop(toks[i], i)), None)
        stop_at = torch.full((1,), steps + 1, dtype=torch.long)
        pos = torch.zeros(1, dtype=torch.long)
        for i in range(steps):
            m._stop_update(stop_at, toks[i], pos)
            pos.add_(1)
        dev = int(stop_at) if int(stop_at) <= steps else None
        assert dev == host, (trial, kind, dev, host)


def test_vectorised_syntax_penalty_matches_reference_loops():
    """sct_gan_b200.syntax (vectorised device scan) against the oracle's restatement of the double loop of
    train.py:334-431, with a fake tokenizer (token -> small id) on id streams dense in the special tokens."""
    from oracle import sct_oracle as O
    from sct_gan_b200.syntax import KEYWORD_FOLLOWERS, SoliditySyntaxRules

    class FakeTok:
        unk_token_id = 3

        def __init__(self):
            words = sorted({w for k, v in KEYWORD_FOLLOWERS.items() for w in [k] + v} | {";", "(", ")", "{", "}"})
            words = [w for w in words if w not in ("interface", "'")]  # two unknown tokens -> unk id
            self.map = {w: 10 + i for i, w in enumerate(words)}

        def convert_tokens_to_ids(self, tok):
            return self.map.get(tok, self.unk_token_id)

    tok = FakeTok()
    V = 64
    rules = SoliditySyntaxRules(tok, V)
    kf = {}
    for kw, fl in KEYWORD_FOLLOWERS.items():
        kid = tok.convert_tokens_to_ids(kw)
        if kid != tok.unk_token_id:
            kf[kid] = [i for i in (tok.convert_tokens_to_ids(f) for f in fl) if i != tok.unk_token_id]
    stmt = [tok.convert_tokens_to_ids(w) for w in ("return", "break", "continue")]
    args = (kf, stmt, tok.map[";"], tok.map["("], tok.map[")"], tok.map["{"], tok.map["}"])
    g = torch.Generator().manual_seed(0)
    for shape in ((2 * 1023,), (3 * 1024,), (2, 300), (1,), (57,)):
        t = torch.randint(0, V, shape, generator=g)
        want = O.syntax_penalty_loops(t, *args)
        got = rules.penalty(t).item()
        assert abs(got - want) < 1e-6, (shape, got, want)
    quiet = torch.full((200,), 5, dtype=torch.long)  # nothing fires -> 0
    assert rules.penalty(quiet).item() == 0.0 and O.syntax_penalty_loops(quiet, *args) == 0.0
