"""The non-differentiable syntax penalty of `SoliditySyntaxLoss._compute_simple_syntax_penalty`
(SCT-GAN/train.py:334-431) as a vectorised integer scan on the device.

The reference walks the shifted target ids in a Python double loop with two `.item()` calls per token (and runs
it twice per step, train.py:947, 953), which caps a training step near one per second whatever the GPU does.
The rules only look at token ids, so they are table look-ups, shifted compares and windowed counts:
  * a keyword followed by a token outside its follower set            -> +2.0   (train.py:383-392)
  * return / break / continue not followed by ';'                     -> +1.5   (train.py:394-398)
  * '(' with no ')' among the next 19 tokens, '{' with no '}' among the next 49 -> +1.0 each (train.py:400-422)
and the result is total / count (0 when nothing fired).  The 2-D reshape rule of train.py:337-350 is kept: a
flat target vector is one sequence unless its length is a multiple of 1024.
"""
from __future__ import annotations

import torch

KEYWORD_FOLLOWERS = {  # train.py:262-282
    "function": ["(", "view", "pure", "external", "public", "internal", "private"],
    "contract": ["{", "is", "interface"],
    "if": ["("], "for": ["("], "while": ["("], "require": ["("], "assert": ["("], "revert": ["("], "emit": ["("],
    "return": [";", "("], "break": [";"], "continue": [";"], "import": ['"', "'"], "pragma": ["solidity"],
    "struct": ["{"], "enum": ["{"], "event": ["("], "modifier": ["{", "("], "mapping": ["("],
}


class SoliditySyntaxRules:
    """Token-id tables built once from a tokenizer exposing `convert_tokens_to_ids` and `unk_token_id`
    (train.py:284-311)."""

    def __init__(self, tokenizer, vocab_size, device="cpu"):
        unk = tokenizer.unk_token_id
        conv = tokenizer.convert_tokens_to_ids
        followers = {}
        for kw, fl in KEYWORD_FOLLOWERS.items():
            kid = conv(kw)
            if kid == unk:
                continue
            ids = [conv(f) for f in fl]
            ids = [i for i in ids if i != unk]
            followers[kid] = ids  # a keyword whose followers are all unknown never fires (train.py:389)
        self.n_kw = len(followers)
        width = max([len(v) for v in followers.values()] + [1])
        self.kw_index = torch.full((vocab_size,), -1, dtype=torch.long)
        self.follow = torch.full((max(self.n_kw, 1), width), -1, dtype=torch.long)
        self.has_follow = torch.zeros(max(self.n_kw, 1), dtype=torch.bool)
        for row, (kid, ids) in enumerate(followers.items()):
            if 0 <= kid < vocab_size:
                self.kw_index[kid] = row
            self.has_follow[row] = len(ids) > 0
            for c, i in enumerate(ids):
                self.follow[row, c] = i
        self.stmt_ids = torch.tensor([conv("return"), conv("break"), conv("continue")], dtype=torch.long)
        self.semicolon, self.lpar, self.rpar = conv(";"), conv("("), conv(")")
        self.lbrace, self.rbrace = conv("{"), conv("}")
        self.to(device)

    def to(self, device):
        for k in ("kw_index", "follow", "has_follow", "stmt_ids"):
            setattr(self, k, getattr(self, k).to(device))
        return self

    @torch.no_grad()
    def penalty(self, target_ids: torch.Tensor) -> torch.Tensor:
        """target_ids: flat [N] (the model's shifted targets) or [B, L].  Returns a 0-dim fp32 tensor."""
        if target_ids.dim() == 1:
            n = target_ids.numel()
            t = target_ids.view(n // 1024, 1024) if n % 1024 == 0 and n > 0 else target_ids.view(1, n)
        else:
            t = target_ids
        B, L = t.shape
        dev = t.device
        if L < 2:
            return torch.zeros((), device=dev)
        cur, nxt = t[:, :-1], t[:, 1:]
        # keyword -> follower set
        row = self.kw_index[cur.clamp(0, self.kw_index.numel() - 1)]
        is_kw = (row >= 0) & (cur >= 0) & (cur < self.kw_index.numel())
        r = row.clamp(min=0)
        ok = (self.follow[r] == nxt.unsqueeze(-1)).any(dim=-1)
        v_kw = is_kw & self.has_follow[r] & ~ok
        # statement keywords need ';'
        v_stmt = (cur.unsqueeze(-1) == self.stmt_ids).any(dim=-1) & (nxt != self.semicolon)

        def unmatched(open_id, close_id, reach):
            c = torch.cumsum((t == close_id).long(), dim=1)  # inclusive counts
            i = torch.arange(L - 1, device=dev)
            j_max = torch.clamp(i + reach - 1, max=L - 1)   # last index of range(i + 1, min(i + reach, L))
            seen = c[:, j_max] - c[:, : L - 1]
            return (cur == open_id) & (seen == 0)

        v_par = unmatched(self.lpar, self.rpar, 20)
        v_brace = unmatched(self.lbrace, self.rbrace, 50)
        total = 2.0 * v_kw.sum() + 1.5 * v_stmt.sum() + 1.0 * v_par.sum() + 1.0 * v_brace.sum()
        count = v_kw.sum() + v_stmt.sum() + v_par.sum() + v_brace.sum()
        return torch.where(count > 0, total / count.clamp(min=1), torch.zeros((), device=dev)).float()
