"""Golden fixture for BASELINE.json configs[0] — the reference's own CPU-runnable case at FULL model size:
default `SmartContractTransformer` (d=768, 8 heads, 6+6 layers, ff=2048, V=50265, max_length=1024, use_gan),
synthetic batch B=8, contract seq S=512, path seq P=128.

    python oracle/make_golden_cfg1.py        (build container only: imports the UNMODIFIED /root/reference)

Runs the reference forward (eval(): dropout off, gradients still flow) + the restated step loss built from the
reference's own loss classes (oracle/make_golden.reference_step) + backward in fp32 on the CPU, checks the oracle
restatement against it, and stores what a [B(T-1), V] = 4088 x 50265 logits tensor and 262.6 M gradients can be
pinned by in about 1.5 MB: every scalar loss, the small outputs in full, per-row logsumexp / argmax / top-2 gap and
256 columns of the logits, the gradient norm of every parameter, full gradients of ten small tensors, the first four
rows of five large ones, the gradient norms of the reference under bf16 autocast (how reproducible the reference
itself is at that precision), and 12 greedy tokens through the reference's own sampling loop.
"""
import contextlib
import io
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/SCT-GAN")

from oracle import sct_oracle as O  # noqa: E402
from oracle.make_golden import build_reference, maxdiff, reference_step  # noqa: E402

B, S, P, SEED = 8, 512, 128, 21


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = dict(O.DEFAULT_CFG)
    t0 = time.time()
    ref = build_reference(cfg)
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    sd = O.synth_state_dict(shapes, SEED)
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    batch = O.make_batch(B, S, P, cfg["vocab_size"], seed=SEED)
    hp = dict(O.DEFAULT_HP)
    golden = {"cfg": cfg, "shapes": {k: list(v) for k, v in shapes.items()}, "seed": SEED,
              "batch_args": dict(B=B, S=S, P=P, vocab=cfg["vocab_size"], seed=SEED), "hp": hp}
    ref.zero_grad(set_to_none=True)
    with contextlib.redirect_stdout(io.StringIO()):
        out, losses = reference_step(ref, batch, hp)
    losses["total_loss"].backward()
    print(f"reference forward + backward: {time.time() - t0:.0f} s")
    # oracle == reference at full size (fp32: six layers deep, so a little looser than the tiny cases)
    with torch.no_grad():
        o_out = O.forward_train(sd, cfg, batch, torch.float32)
        o_loss = O.step_losses(o_out, batch, hp)
    for k in ("logits", "contract_vulnerability_logits", "line_vulnerability_logits", "encoder_output",
              "discriminator_logits"):
        dlt = maxdiff(out[k], o_out[k])
        assert dlt <= 1e-3 * max(1.0, out[k].abs().max().item()), (k, dlt)
    assert torch.equal(out["target_ids"], o_out["target_ids"])
    for k in ("gen_ce_loss", "contract_vuln_loss", "line_vuln_loss", "discriminator_loss", "total_loss"):
        assert abs(float(losses[k]) - float(o_loss[k])) <= 1e-4 * max(1.0, abs(float(losses[k]))), k
    print("cfg1: oracle == reference (fp32)")
    logits = out["logits"].detach()
    top2 = logits.topk(2, dim=-1)
    g = torch.Generator().manual_seed(5)
    cols = torch.cat([torch.arange(128), torch.randperm(cfg["vocab_size"] - 128, generator=g)[:128] + 128])
    n_lines = int(batch["token_to_line"].max()) + 1
    golden["outputs"] = {
        "target_ids": out["target_ids"].clone(),
        "encoder_output": out["encoder_output"].detach().clone(),
        "contract_vulnerability_logits": out["contract_vulnerability_logits"].detach().clone(),
        "discriminator_logits": out["discriminator_logits"].detach().clone(),
        "line_vulnerability_logits": out["line_vulnerability_logits"].detach()[:, :n_lines].clone(),
        "logits_lse": torch.logsumexp(logits, dim=-1), "logits_argmax": top2.indices[:, 0].clone(),
        "logits_top2_gap": (top2.values[:, 0] - top2.values[:, 1]).clone(),
        "logits_cols": cols, "logits_at_cols": logits[:, cols].half(),
    }
    golden["losses"] = {k: float(v) for k, v in losses.items()}
    grads = {n: p.grad for n, p in ref.named_parameters()}
    golden["grad_norms"] = {n: (float(gr.norm()) if gr is not None else None) for n, gr in grads.items()}
    keep = ["output_norm.weight", "output_layer.bias", "embedding_norm.weight", "ast_embedding_norm.bias",
            "encoder.layers.0.norm1.weight", "decoder.layers.5.norm2.bias", "feature_fusion.8.bias",
            "disc_synthetic_head.4.weight", "decoder.layers.3.multihead_attn.in_proj_bias",
            "encoder.layers.5.linear1.bias"]
    golden["grads"] = {n: grads[n].detach().clone() for n in keep}
    golden["grad_rows"] = {n: grads[n][:4].detach().clone() for n in
                           ("output_layer.weight", "encoder.layers.0.self_attn.in_proj_weight",
                            "decoder.layers.5.linear2.weight", "ast_attention.out_proj.weight",
                            "decoder.layers.2.multihead_attn.in_proj_weight")}
    # rows of the embedding gradients that the batch touches (ids >= 3: the first four rows are never used)
    ids = batch["input_ids"][0, :4]
    golden["embedding_grad_ids"] = ids.clone()
    golden["embedding_grad_rows"] = grads["embedding.weight"][ids].detach().clone()
    ref.zero_grad(set_to_none=True)
    with contextlib.redirect_stdout(io.StringIO()), torch.autocast("cpu", dtype=torch.bfloat16):
        _, losses_ac = reference_step(ref, batch, hp)
    losses_ac["total_loss"].float().backward()
    golden["grad_norms_autocast"] = {n: (float(p.grad.float().norm()) if p.grad is not None else None)
                                     for n, p in ref.named_parameters()}
    golden["losses_autocast"] = {k: float(v) for k, v in losses_ac.items()}
    print(f"autocast pass done: {time.time() - t0:.0f} s")
    n_new = 12
    with torch.no_grad():
        toks, gaps = O.generate_greedy(sd, cfg, batch, n_new)
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
    print(f"cfg1: greedy generation oracle == reference ({n_new} tokens)")
    golden["greedy_tokens"] = toks
    golden["greedy_gaps"] = gaps
    path = os.path.join(ROOT, "tests", "golden_cfg1", "cfg1_default_model.pt")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(golden, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", f"({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
