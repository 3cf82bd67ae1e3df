"""Generates tests/golden/*.pt from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py

Imports /root/reference/SCT-GAN/model.py (SmartContractTransformer) and train.py (loss classes), loads the
deterministic synthetic weights of oracle.sct_oracle.synth_state_dict, runs the reference forward (eval(),
dropout off) + the restated step loss built from the reference's own loss classes + backward, and
  1. checks the oracle restatement against the reference (fp64 <= 1e-9, fp32 <= 2e-4),
  2. writes the reference's outputs as small fixtures that travel to the GPU box.
/root/reference does not exist on the GPU box; nothing outside this script reads it.
"""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/SCT-GAN"
sys.path.insert(0, REF)

from oracle import sct_oracle as O  # noqa: E402

CASES = {
    # name: (cfg overrides, B, S, P, seed)
    "tiny_ragged": (dict(num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=512, max_length=128,
                         vocab_size=1000), 3, 48, 40, 11),
    "tiny_len130": (dict(num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=256, max_length=256,
                         vocab_size=777), 2, 130, 33, 12),
}


def build_reference(cfg):
    with contextlib.redirect_stdout(io.StringIO()):  # the reference prints DEBUG lines while initialising
        from model import SmartContractTransformer as Ref
        m = Ref(**cfg)
    return m


def reference_step(model, batch, hp):
    """train.py:917-924 forward + the loss arithmetic of :937-997, :1185-1270 using the reference's classes."""
    import train as T

    out = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                ast_input_ids=batch["ast_input_ids"], ast_attention_mask=batch["ast_attention_mask"],
                target_ids=batch["target_ids"], token_to_line=batch["token_to_line"])
    dt = out["logits"].dtype
    ce = torch.nn.functional.cross_entropy(out["logits"], out["target_ids"], reduction="mean")  # train.py:324
    gen = ce + 0.5 * hp["syntax_penalty"]
    cfl = T.ContractLevelFocalLoss(alpha=0.05, gamma=4.0, reduction="mean")
    sfl = T.SpatialAwareFocalLoss(alpha=0.25, gamma=2.0, spatial_weight=0.2, reduction="mean")
    cv = cfl(out["contract_vulnerability_logits"], batch["contract_vulnerabilities"].to(dt))
    lv = sfl(out["line_vulnerability_logits"].view(-1, 8), batch["vulnerable_lines"].view(-1, 8).to(dt),
             batch["token_to_line"].view(-1))
    cv = torch.max(cv, torch.tensor(0.0001, dtype=dt))
    lv = torch.max(lv, torch.tensor(0.000001, dtype=dt))
    if lv > 5.0:
        lv = lv * 0.1
    elif lv > 1.0:
        lv = lv * 0.5
    w_line = hp["line_vuln_weight"] * hp["warmup_factor"] * hp["stability_factor"] * hp["line_loss_scale"]
    bce = torch.nn.BCEWithLogitsLoss()
    z = out["discriminator_logits"]
    d_loss = bce(z, torch.ones_like(z))
    conf = torch.sigmoid(z).mean().item()
    adv = torch.zeros((), dtype=dt)
    if conf < 0.3:
        adv = bce(z, torch.zeros_like(z))
    if conf > 0.8:
        d_loss = d_loss + 1.0 * torch.mean(torch.sigmoid(z) ** 2) + 2.0 * torch.mean(torch.sigmoid(z) ** 4)
    total = 0.5 * gen + 0.25 * cv * hp["contract_vuln_weight"] + 0.2 * lv * w_line + 0.05 * d_loss
    if adv > 0:
        total = total + 0.02 * adv
    losses = dict(gen_ce_loss=ce, contract_vuln_loss=cv, line_vuln_loss=lv, discriminator_loss=d_loss,
                  adversarial_loss=adv, discriminator_confidence=conf, total_loss=total)
    return out, losses


def maxdiff(a, b):
    return (a.double() - b.double()).abs().max().item()


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    torch.manual_seed(0)
    for name, (over, B, S, P, seed) in CASES.items():
        cfg = {**O.DEFAULT_CFG, **over}
        ref = build_reference(cfg)
        shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        sd = O.synth_state_dict(shapes, seed)
        ref.load_state_dict(sd, strict=True)
        ref.eval()
        batch = O.make_batch(B, S, P, cfg["vocab_size"], seed=seed)
        hp = dict(O.DEFAULT_HP)
        golden = {"cfg": cfg, "shapes": {k: list(v) for k, v in shapes.items()}, "seed": seed,
                  "batch_args": dict(B=B, S=S, P=P, vocab=cfg["vocab_size"], seed=seed), "hp": hp}
        for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 2e-4)):
            ref.to(dtype)
            ref.zero_grad(set_to_none=True)
            with contextlib.redirect_stdout(io.StringIO()):
                out, losses = reference_step(ref, batch, hp)
            losses["total_loss"].backward()
            sd_t = {k: v.to(dtype) for k, v in sd.items()}
            o_out = O.forward_train(sd_t, cfg, batch, dtype)
            o_loss = O.step_losses(o_out, batch, hp)
            for k in ("logits", "contract_vulnerability_logits", "line_vulnerability_logits", "encoder_output",
                      "discriminator_logits"):
                dlt = maxdiff(out[k], o_out[k])
                assert dlt <= tol * max(1.0, out[k].abs().max().item()), (name, dtype, k, dlt)
            assert torch.equal(out["target_ids"], o_out["target_ids"])
            for k in ("gen_ce_loss", "contract_vuln_loss", "line_vuln_loss", "discriminator_loss", "total_loss"):
                dlt = abs(float(losses[k]) - float(o_loss[k]))
                assert dlt <= tol * max(1.0, abs(float(losses[k]))), (name, dtype, k, dlt)
            print(f"{name} {dtype}: oracle == reference (tol {tol})")
            if dtype == torch.float32:
                golden["outputs"] = {k: out[k].detach().clone() for k in
                                     ("contract_vulnerability_logits", "encoder_output", "discriminator_logits")}
                golden["outputs"]["logits"] = out["logits"].detach().clone()
                golden["outputs"]["target_ids"] = out["target_ids"].clone()
                lvl = out["line_vulnerability_logits"].detach()
                n_lines = int(batch["token_to_line"].max()) + 1
                golden["outputs"]["line_vulnerability_logits"] = lvl[:, :n_lines].clone()
                golden["losses"] = {k: float(v) for k, v in losses.items()}
                grads = {n: p.grad for n, p in ref.named_parameters()}
                golden["grad_norms"] = {n: (float(g.norm()) if g is not None else None) for n, g in grads.items()}
                keep = ["output_norm.weight", "output_layer.bias", "embedding_norm.weight", "ast_embedding_norm.bias",
                        "encoder.layers.0.norm1.weight", "decoder.layers.0.norm2.bias", "feature_fusion.8.bias",
                        "disc_synthetic_head.4.weight", "decoder.layers.0.multihead_attn.in_proj_bias",
                        "encoder.layers.0.linear1.bias"]
                golden["grads"] = {n: grads[n].detach().clone() for n in keep}
                # rows of big matrix gradients (first 4 rows) as a cheap fingerprint
                golden["grad_rows"] = {n: grads[n][:4].detach().clone() for n in
                                       ("output_layer.weight", "embedding.weight", "encoder.layers.0.self_attn.in_proj_weight",
                                        "decoder.layers.0.linear2.weight", "ast_attention.out_proj.weight")}
        # the reference itself under bf16 autocast: how reproducible its gradients are at bf16 precision
        ref.to(torch.float32)
        ref.zero_grad(set_to_none=True)
        with contextlib.redirect_stdout(io.StringIO()), torch.autocast("cpu", dtype=torch.bfloat16):
            _, losses_ac = reference_step(ref, batch, hp)
        losses_ac["total_loss"].float().backward()
        golden["grad_norms_autocast"] = {n: (float(p.grad.float().norm()) if p.grad is not None else None)
                                         for n, p in ref.named_parameters()}
        golden["losses_autocast"] = {k: float(v) for k, v in losses_ac.items()}
        # greedy generation through the reference's own loop (multinomial -> argmax), fp32
        ref.to(torch.float32)
        n_new = 12
        sd32 = {k: v.float() for k, v in sd.items()}
        toks, gaps = O.generate_greedy(sd32, cfg, batch, n_new)
        real_multinomial = torch.multinomial
        torch.multinomial = lambda probs, num_samples: probs.argmax(dim=-1, keepdim=True)
        ref.max_length = n_new + 1
        try:
            with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
                gen = ref(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                          ast_input_ids=batch["ast_input_ids"], ast_attention_mask=batch["ast_attention_mask"],
                          target_ids=None, token_to_line=batch["token_to_line"])
        finally:
            torch.multinomial = real_multinomial
        assert torch.equal(gen["generated_sequence"], toks), (gen["generated_sequence"], toks)
        print(f"{name}: greedy generation oracle == reference ({n_new} tokens)")
        golden["greedy_tokens"] = toks
        golden["greedy_gaps"] = gaps
        torch.save(golden, os.path.join(ROOT, "tests", "golden", f"{name}.pt"))
        print("wrote", name, os.path.getsize(os.path.join(ROOT, "tests", "golden", f"{name}.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
