"""§8f-3 / §8f-4 on the GPU: the device versions of the reference's loss-side host loops and of the sampling tail,
run on CUDA tensors, against the oracle's literal restatements of the reference loops (CPU).

  * spatial penalty of SpatialAwareFocalLoss (train.py:174-245) — live only at S == 1024
  * SoliditySyntaxLoss penalty scan (train.py:334-431)
  * adaptive-threshold line metrics (train.py:1043-1140)
  * temperature / top-k / top-p / multinomial tail of the sampling loop (model.py:892-918): chi-square test of the
    sampled distribution against the distribution those lines define
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_spatial_penalty_s1024_on_gpu_matches_reference_loop(cuda_dev):
    from oracle import sct_oracle as O
    from sct_gan_b200.trainer import spatial_aware_focal_loss, spatial_penalty

    g = torch.Generator().manual_seed(4)
    B, S, C = 2, 1024, 8  # token_to_line.numel() == B * 1024 rows of pred: the only case in which the penalty is live
    pred = torch.randn(B * S, C, generator=g)
    target = (torch.rand(B * S, C, generator=g) < 0.01).float()
    t2l = (torch.arange(S) // 12)[None, :].expand(B, S).reshape(-1).contiguous()
    want = O.spatial_penalty(pred, target, t2l)
    assert float(want.abs().max()) > 0  # the case is live
    for n_lines in (None, 86):
        got = spatial_penalty(pred.cuda(), target.cuda(), t2l.cuda(), n_lines)
        assert (got.cpu() - want).abs().max().item() < 1e-6
    # ... and differentiable through the in-place writes of train.py:236-239: gradient of the whole focal loss
    pc = pred.clone().requires_grad_(True)
    O.spatial_focal_loss(pc, target, t2l, 0.25, 2.0, 0.2).backward()
    pg = pred.cuda().requires_grad_(True)
    a = torch.tensor([0.25, 2.0, 0.2], device="cuda")
    spatial_aware_focal_loss(pg, target.cuda(), t2l.cuda(), a[0], a[1], a[2], 86).backward()
    assert (pg.grad.cpu() - pc.grad).abs().max().item() < 1e-7 + 1e-4 * pc.grad.abs().max().item()
    # any other S: the reference silently returns zeros
    short = spatial_penalty(pred[:700].cuda(), target[:700].cuda(), t2l[:512].cuda())
    assert float(short.abs().max()) == 0.0


def test_syntax_penalty_on_gpu_matches_reference_loops(cuda_dev):
    from oracle import sct_oracle as O
    from sct_gan_b200.syntax import KEYWORD_FOLLOWERS, SoliditySyntaxRules

    class FakeTok:
        unk_token_id = 3

        def __init__(self):
            words = sorted({w for k, v in KEYWORD_FOLLOWERS.items() for w in [k] + v} | {";", "(", ")", "{", "}"})
            words = [w for w in words if w not in ("interface", "'")]
            self.map = {w: 10 + i for i, w in enumerate(words)}

        def convert_tokens_to_ids(self, tok):
            return self.map.get(tok, self.unk_token_id)

    tok = FakeTok()
    V = 64
    rules = SoliditySyntaxRules(tok, V, "cuda")
    kf = {}
    for kw, fl in KEYWORD_FOLLOWERS.items():
        kid = tok.convert_tokens_to_ids(kw)
        if kid != tok.unk_token_id:
            kf[kid] = [i for i in (tok.convert_tokens_to_ids(f) for f in fl) if i != tok.unk_token_id]
    stmt = [tok.convert_tokens_to_ids(w) for w in ("return", "break", "continue")]
    args = (kf, stmt, tok.map[";"], tok.map["("], tok.map[")"], tok.map["{"], tok.map["}"])
    g = torch.Generator().manual_seed(0)
    for shape in ((2 * 1023,), (3 * 1024,), (2, 300), (57,), (32 * 1023,)):
        t = torch.randint(0, V, shape, generator=g)
        if t.numel() > 8000:  # the B = 32 shape of the bench step: the loop oracle is O(N * 50), check a slice consistently
            t = t[: 4 * 1023]
        want = O.syntax_penalty_loops(t, *args)
        got = rules.penalty(t.cuda()).item()
        assert abs(got - want) < 1e-6, (shape, got, want)


def test_line_metrics_on_gpu_match_reference_cascade(cuda_dev):
    from oracle import sct_oracle as O
    from sct_gan_b200.trainer import line_vulnerability_metrics

    g = torch.Generator().manual_seed(11)
    B, L, T = 4, 1024, 8
    cases = {
        "ordinary": torch.randn(B, L, T, generator=g),
        "negative logits": torch.randn(B, L, T, generator=g) - 3.0,
        "too many": torch.randn(B, L, T, generator=g) * 0.05 + 2.0,
        "fallback": torch.full((B, L, T), -0.9) + 0.01 * torch.randn(B, L, T, generator=g),
        "ultra fallback": torch.full((B, L, T), -6.0) + 0.01 * torch.randn(B, L, T, generator=g),
    }
    for name, logits in cases.items():
        vl = (torch.rand(B, L, T, generator=g) < 0.02).float()
        ref = O.line_metrics_loops(logits, vl)
        got = line_vulnerability_metrics(logits.cuda(), vl.cuda())
        mine = (got["line_vuln_accuracy"].item(), got["line_vuln_precision"].item(), got["line_vuln_recall"].item(),
                got["line_vuln_threshold"].item(), float(got["line_vuln_predictions"].item()))
        for a_, b_ in zip(mine, ref):
            # the GPU quantile interpolates in fp32 like the CPU one; thresholds sit between samples, counts must agree
            assert abs(a_ - b_) < 1e-5 * max(1.0, abs(b_)), (name, mine, ref)


def test_sampling_tail_distribution_matches_reference_rule(cuda_dev):
    """model._sample (non-greedy) on fixed logits: N draws per row against the distribution model.py:892-918 defines
    (logits / 0.7 -> keep the 50 largest -> softmax -> drop the tail beyond cumulative 0.95, always keeping the first
    -> renormalise -> multinomial).  Chi-square with the usual 5-expected-counts pooling, alpha = 1e-4 per row."""
    from sct_gan_b200 import SmartContractTransformer

    m = SmartContractTransformer.__new__(SmartContractTransformer)  # only the sampling method is used
    g = torch.Generator().manual_seed(9)
    V, N = 5000, 40000
    for trial, spread in enumerate((1.0, 3.0, 0.2)):
        # bf16-representable logits (the sampler reads the vocab GEMM's bf16 rows); such values tie, and ties resolve
        # to the lowest index: a stable descending sort states the same rule
        row = (torch.randn(V, generator=g) * spread).bfloat16().float()
        # the reference rule, written out on the CPU in fp64
        z = row.double() / 0.7
        srt = torch.sort(z, descending=True, stable=True)
        topv, topi = srt.values[:50], srt.indices[:50]
        pr = torch.softmax(topv, dim=-1)
        remove = torch.cumsum(pr, dim=-1) > 0.95
        remove[1:] = remove[:-1].clone()
        remove[0] = False
        pr = torch.softmax(topv.masked_fill(remove, float("-inf")), dim=-1)
        expected = torch.zeros(V, dtype=torch.float64)
        expected[topi] = pr
        logits = row.cuda().to(torch.bfloat16)[None, :].expand(N, V).contiguous()
        prev = torch.ones((N, 1), dtype=torch.long, device="cuda")
        torch.manual_seed(100 + trial)
        draws = m._sample(logits, prev, apply_syntax_constraints=False, greedy=False).view(-1).cpu()
        counts = torch.bincount(draws, minlength=V).double()
        assert float(counts[expected == 0].sum()) == 0.0  # nothing outside the nucleus is ever drawn
        exp_n = expected * N
        big = exp_n >= 5.0
        chi2 = float((((counts - exp_n) ** 2) / exp_n.clamp(min=1e-30))[big].sum())
        rest_e, rest_o = float(exp_n[~big].sum()), float(counts[~big].sum())
        dof = int(big.sum()) - 1
        if rest_e >= 5.0:
            chi2 += (rest_o - rest_e) ** 2 / rest_e
            dof += 1
        # Wilson-Hilferty bound on the chi-square quantile at alpha = 1e-4 (z = 3.72)
        bound = dof * (1 - 2 / (9 * dof) + 3.72 * math.sqrt(2 / (9 * dof))) ** 3 if dof > 0 else 0.0
        assert chi2 <= bound or dof == 0, (trial, chi2, dof, bound)
        if dof == 0:  # the nucleus collapsed to one token: every draw is that token
            assert float(counts[topi[0]]) == N
    # the syntax tweak of model.py:975-1060 (';' logit doubled after ids 2000-2002) moves mass only in those rows
    V2 = 3000
    lg = torch.zeros(4, V2, device="cuda")
    lg[:, 59] = 1.0
    prev = torch.tensor([[2001], [5], [2000], [2003]], device="cuda")
    out = m._apply_syntax_constraints(lg, prev)
    assert out[:, 59].tolist() == [2.0, 1.0, 2.0, 1.0]
