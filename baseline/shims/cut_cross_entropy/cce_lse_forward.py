from __future__ import annotations

from dataclasses import dataclass

import torch

from ._chunked import chunk_logits, row_chunks


@dataclass
class LSEReturn:
    lse: torch.Tensor
    neg_correct_logit: torch.Tensor | None
    logit_avg: torch.Tensor | None


def cce_lse_forward_kernel(e: torch.Tensor, c: torch.Tensor, bias: torch.Tensor | None = None, valids: torch.Tensor | None = None,
                           softcap: float | None = None, return_logit_avg: bool = False, shift: int = 0,
                           targets: torch.Tensor | None = None) -> LSEReturn:
    rows = e if valids is None else e.index_select(0, valids.long())
    n = rows.shape[0]
    tgt = None
    if targets is not None:
        tgt = targets if valids is None else targets.index_select(0, (valids + shift).long())
    lse = torch.empty(n, dtype=torch.float32, device=e.device)
    neg = torch.empty(n, dtype=torch.float32, device=e.device) if tgt is not None else None
    step = row_chunks(n, c.shape[0])
    for s in range(0, n, step):
        logits, _ = chunk_logits(rows[s : s + step], c, bias, softcap)
        lse[s : s + step] = torch.logsumexp(logits, dim=-1)
        if tgt is not None:
            t = tgt[s : s + step].long()
            ok = (t >= 0) & (t < c.shape[0])  # vocab-parallel: targets owned by another rank are out of range
            got = logits.gather(1, t.clamp(0, c.shape[0] - 1)[:, None]).squeeze(1)
            neg[s : s + step] = torch.where(ok, -got, torch.zeros_like(got))
    # logit_avg only feeds the vocabulary re-ordering of the real backward kernel's gradient filter; the stand-in
    # backward is exact, so it is not produced
    return LSEReturn(lse=lse, neg_correct_logit=neg, logit_avg=None)
