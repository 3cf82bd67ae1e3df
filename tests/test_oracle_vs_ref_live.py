"""The oracle restatement against the UNMODIFIED reference, live: `oracle/_ref` (the reference's model.py / train.py
byte-compiled by oracle/build_ref.py, loaded by oracle/ref_loader.py) is run next to `oracle.sct_oracle` on the same
synthetic weights and batch.  Skipped where oracle/_ref has not been built (it needs /root/reference once); the
committed fixtures under tests/golden/ carry the same comparison to machines without it."""
import contextlib
import io

import pytest
import torch

from oracle import ref_loader
from oracle import sct_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not built")

CFG = dict(num_encoder_layers=1, num_decoder_layers=2, dim_feedforward=256, max_length=128, vocab_size=611)


def _reference(cfg, seed):
    model_mod, _ = ref_loader.load()
    with contextlib.redirect_stdout(io.StringIO()):  # the reference prints while initialising
        ref = model_mod.SmartContractTransformer(**cfg)
    sd = O.synth_state_dict({k: tuple(v.shape) for k, v in ref.state_dict().items()}, seed)
    ref.load_state_dict(sd, strict=True)
    return ref.eval(), sd


def test_forward_matches_reference_fp64():
    cfg = {**O.DEFAULT_CFG, **CFG}
    ref, sd = _reference(cfg, 21)
    batch = O.make_batch(3, 50, 37, cfg["vocab_size"], seed=21)
    ref.to(torch.float64)
    with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
        out = ref(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                  ast_input_ids=batch["ast_input_ids"], ast_attention_mask=batch["ast_attention_mask"],
                  target_ids=batch["target_ids"], token_to_line=batch["token_to_line"])
    o = O.forward_train({k: v.double() for k, v in sd.items()}, cfg, batch, torch.float64)
    assert torch.equal(out["target_ids"], o["target_ids"])  # shifted targets: index work, bit-exact
    for k in ("logits", "encoder_output", "contract_vulnerability_logits", "line_vulnerability_logits",
              "discriminator_logits"):
        d = (out[k].double() - o[k].double()).abs().max().item()
        assert d <= 1e-9 * max(1.0, out[k].abs().max().item()), (k, d)


def test_loss_classes_match_reference_fp64():
    _, train_mod = ref_loader.load()
    g = torch.Generator().manual_seed(3)
    B, S, C = 3, 1024, 8  # S == 1024: the spatial penalty's line branch is live (train.py:174-245)
    logits = torch.randn(B, S, C, generator=g, dtype=torch.float64)
    tgt = (torch.rand(B, S, C, generator=g) < 0.02).double()
    t2l = (torch.arange(S) // 9)[None].expand(B, S).contiguous()
    sfl = train_mod.SpatialAwareFocalLoss(alpha=0.25, gamma=2.0, spatial_weight=0.2, reduction="mean")
    with contextlib.redirect_stdout(io.StringIO()):
        want = sfl(logits.view(-1, C), tgt.view(-1, C), t2l.reshape(-1))
    got = O.spatial_focal_loss(logits.view(-1, C), tgt.view(-1, C), t2l.reshape(-1), 0.25, 2.0, 0.2)
    assert abs(float(want) - float(got)) <= 1e-9 * max(1.0, abs(float(want)))
    cl = torch.randn(B, C, generator=g, dtype=torch.float64)
    ct = (torch.rand(B, C, generator=g) < 0.3).double()
    cfl = train_mod.ContractLevelFocalLoss(alpha=0.05, gamma=4.0, reduction="mean")
    assert abs(float(cfl(cl, ct)) - float(O.contract_focal_loss(cl, ct, 0.05, 4.0))) <= 1e-12
