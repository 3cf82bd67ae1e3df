"""One adversarial train step of the UNMODIFIED reference on CPU — the CPU arm of bench.py.

The model is the reference's own `SmartContractTransformer` (oracle/_ref/model.refbin, byte-compiled from
/root/reference/SCT-GAN/model.py by oracle/build_ref.py), the losses are the reference's own classes
(train.py: ContractLevelFocalLoss, SpatialAwareFocalLoss), the optimiser is torch.optim.AdamW with the groups of
train.py:512-540.  `SmartContractTrainer` itself cannot be constructed (ReduceLROnPlateau(verbose=True) is a
TypeError on torch 2.11, the device is hard-wired to CUDA, the syntax loss needs a tokenizer download), so the body
of its batch loop is restated here line by line:
  train.py:909-924   forward (target = augmented target ids), train mode: dropout 0.3 active
  train.py:937-947   generator CE (F.cross_entropy mean, :324); the syntax penalty is a constant (0 here)
  train.py:974-997   contract / line focal losses with the reference's classes
  train.py:1185-1194 floors and the > 1 / > 5 rescale
  train.py:1201-1234 GAN terms with the 0.3 / 0.8 confidence branches (host .item(), as the reference does)
  train.py:1245-1270 loss weights (use_augmentation and use_gan branch)
  train.py:1273-1311 zero_grad, backward, three clips, the per-parameter .item() norm loop, skip rules, AdamW
Test / measurement infrastructure only."""
import contextlib
import io
import os
import time

import torch

from . import ref_loader


class ReferenceStep:
    def __init__(self, max_length=1024, learning_rate=1e-6, weight_decay=0.1, seed=0, threads=None, **cfg):
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.M, self.T = ref_loader.load()
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):  # the reference prints while initialising
            self.model = self.M.SmartContractTransformer(use_gan=True, max_length=max_length, **cfg)
        # the reference zero-initialises every 1-D parameter (all logits 0): re-draw them (SURVEY §7 hard part 1)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for n, p in self.model.named_parameters():
                if p.dim() == 1:
                    noise = torch.randn(p.shape, generator=g)
                    p.copy_(1.0 + 0.1 * noise if n.endswith("weight") else 0.02 * noise)
        self.model.train()
        base, contract, line, disc = [], [], [], []
        for name, p in self.model.named_parameters():  # train.py:518-527
            if "disc_" in name:
                disc.append(p)
            elif ("contract_vulnerability_head" in name or "contract_feature_aggregation" in name
                  or "contract_vuln_attention" in name):
                contract.append(p)
            elif ("line_vulnerability_head" in name or "line_feature_extractor" in name
                  or "line_vuln_attention" in name or "vuln_type_attention" in name):
                line.append(p)
            else:
                base.append(p)
        groups = [{"params": base, "lr": learning_rate}, {"params": contract, "lr": learning_rate * 2.0},
                  {"params": line, "lr": learning_rate * 3.0}, {"params": disc, "lr": learning_rate * 0.5}]
        self.optimizer = torch.optim.AdamW(groups, weight_decay=weight_decay, betas=(0.9, 0.98), eps=1e-9)
        self.cfl = self.T.ContractLevelFocalLoss(alpha=0.05, gamma=4.0, reduction="mean")           # train.py:561-565
        self.sfl = self.T.SpatialAwareFocalLoss(alpha=0.25, gamma=2.0, spatial_weight=0.2, reduction="mean")  # :568-573
        self.bce = torch.nn.BCEWithLogitsLoss()                                                       # train.py:612
        self.max_grad_norm = 1.0
        self.line_vuln_weight, self.contract_vuln_weight = 2.0, 3.0
        self.warmup = 1.0 / 5.0  # epoch 0 of 5 warm-up epochs (train.py:906-907)

    def step(self, batch):
        m = self.model
        with contextlib.redirect_stdout(io.StringIO()):
            out = m(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                    ast_input_ids=batch["ast_input_ids"], ast_attention_mask=batch["ast_attention_mask"],
                    target_ids=batch["target_ids"], token_to_line=batch["token_to_line"])
            gen = torch.nn.functional.cross_entropy(out["logits"], out["target_ids"], reduction="mean")
            cv = self.cfl(out["contract_vulnerability_logits"], batch["contract_vulnerabilities"].float())
            lv = self.sfl(out["line_vulnerability_logits"].view(-1, 8), batch["vulnerable_lines"].view(-1, 8).float(),
                          batch["token_to_line"].reshape(-1))
        cv = torch.max(cv, torch.tensor(0.0001))
        lv = torch.max(lv, torch.tensor(0.000001))
        if lv > 5.0:
            lv = lv * 0.1
        elif lv > 1.0:
            lv = lv * 0.5
        z = out["discriminator_logits"]
        d_loss = self.bce(z, torch.ones_like(z))
        conf = torch.sigmoid(z).mean().item()
        adv = 0.0
        if conf < 0.3:
            adv = self.bce(z, torch.zeros_like(z))
        if conf > 0.8:
            d_loss = d_loss + 1.0 * torch.mean(torch.sigmoid(z) ** 2) + 2.0 * torch.mean(torch.sigmoid(z) ** 4)
        w_line = self.line_vuln_weight * self.warmup
        total = 0.5 * gen + 0.25 * cv * self.contract_vuln_weight + 0.2 * lv * w_line + 0.05 * d_loss
        if torch.is_tensor(adv):
            total = total + 0.02 * adv
        self.optimizer.zero_grad()
        total.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), self.max_grad_norm)
        dp = [p for n, p in m.named_parameters() if "disc_" in n]
        torch.nn.utils.clip_grad_norm_(dp, self.max_grad_norm * 0.3)
        vp = [p for n, p in m.named_parameters() if "vulnerability_head" in n or "line_feature_extractor" in n
              or "line_vuln_attention" in n or "vuln_type_attention" in n]
        torch.nn.utils.clip_grad_norm_(vp, self.max_grad_norm * 2.0)
        total_norm = 0.0
        for p in m.parameters():
            if p.grad is not None:
                total_norm += p.grad.data.norm(2).item() ** 2
        total_norm = total_norm ** 0.5
        if torch.isnan(total) or torch.isinf(total) or total_norm > 1000:
            self.optimizer.zero_grad()
            return float(total.detach()), total_norm, False
        self.optimizer.step()
        return float(total.detach()), total_norm, True


def synthetic_batch(B, S, P, vocab, seed, lines_per=12):
    """Same construction as bench.synthetic_batch."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, vocab, (B, S), generator=g)
    ast = torch.randint(3, vocab, (B, P), generator=g)
    tgt = torch.randint(3, vocab, (B, S), generator=g)
    ls = torch.randint(S // 2, S + 1, (B,), generator=g)
    lp = torch.randint(max(1, P // 2), P + 1, (B,), generator=g)
    return dict(
        input_ids=ids, attention_mask=(torch.arange(S)[None, :] < ls[:, None]).long(),
        ast_input_ids=ast, ast_attention_mask=(torch.arange(P)[None, :] < lp[:, None]).long(),
        target_ids=tgt, token_to_line=(torch.arange(S) // lines_per)[None, :].expand(B, S).contiguous(),
        contract_vulnerabilities=(torch.rand(B, 8, generator=g) < 0.2).float(),
        vulnerable_lines=(torch.rand(B, 1024, 8, generator=g) < 0.01).float())


def time_reference(B, S, P, steps, warmup, max_length=1024, seed=1234):
    """Returns (seconds per step, tokens per step, threads, last total loss)."""
    rs = ReferenceStep(max_length=max(max_length, S))
    batch = synthetic_batch(B, S, P, 50265, seed)
    times, loss = [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loss, _, _ = rs.step(batch)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), B * S, rs.threads, loss
