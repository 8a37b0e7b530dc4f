from __future__ import annotations

import torch
import torch.distributed as dist

from ._chunked import chunk_logits, row_chunks


def cce_backward_kernel(do: torch.Tensor, dlse: torch.Tensor | None, e: torch.Tensor, e_info, c: torch.Tensor, c_info,
                        bias: torch.Tensor | None, bias_info, lse: torch.Tensor, valids: torch.Tensor | None,
                        softcap: float | None, filter_eps: float | None, targets: torch.Tensor, shift: int = 0,
                        vocab_ordering: torch.Tensor | None = None, grad_scale: float = 1.0, accum_e_fp32: bool = False,
                        accum_c_fp32: bool = False, filter_e_grad: bool = True, filter_c_grad: bool = True,
                        reduce_e_grad: bool = False, pg=None):
    """Exact gradients of ``sum_i do_i * nll_i + dlse_i * lse_i`` wrt ``e``, ``c`` and ``bias`` (row-chunked)."""
    vidx = None if valids is None else valids.long()
    rows = e if vidx is None else e.index_select(0, vidx)
    n, vocab = rows.shape[0], c.shape[0]
    tidx = None if vidx is None else (valids + shift).long()
    tgt = targets if tidx is None else targets.index_select(0, tidx)
    if do.numel() == 1:
        g_nll = do.reshape(1).float().expand(n)
    else:
        g_nll = (do if tidx is None else do.index_select(0, tidx)).float()
    g_lse = None
    if dlse is not None:
        g_lse = (dlse if tidx is None else dlse.index_select(0, tidx)).float()
    need_e = e_info.requires_grad
    need_c = c_info.requires_grad
    need_b = bias is not None and bias_info is not None and bias_info.requires_grad
    de_rows = torch.empty_like(rows) if need_e else None
    dc = torch.zeros(c.shape, dtype=torch.float32, device=c.device) if need_c else None
    db = torch.zeros(vocab, dtype=torch.float32, device=c.device) if need_b else None
    step = row_chunks(n, vocab)
    for s in range(0, n, step):
        er = rows[s : s + step]
        logits, tanh_term = chunk_logits(er, c, bias, softcap)
        p = torch.exp(logits.sub_(lse[s : s + step, None]))
        coef = g_nll[s : s + step]
        if g_lse is not None:
            coef = coef + g_lse[s : s + step]
        d = p.mul_(coef[:, None])
        t = tgt[s : s + step].long()
        ok = (t >= 0) & (t < vocab)
        d.scatter_add_(1, t.clamp(0, vocab - 1)[:, None], torch.where(ok, -g_nll[s : s + step], torch.zeros_like(coef))[:, None])
        if tanh_term is not None:
            d.mul_(1 - tanh_term * tanh_term)
        if grad_scale != 1.0:
            d.mul_(grad_scale)
        if need_b:
            db += d.sum(0)
        d16 = d.to(e.dtype)
        if need_e:
            torch.mm(d16, c, out=de_rows[s : s + step])
        if need_c:
            dc += torch.mm(d16.t(), er)
    de = None
    if need_e:
        if vidx is None:
            de = de_rows
        else:
            de = torch.zeros_like(e)
            de.index_copy_(0, vidx, de_rows)
        if reduce_e_grad and pg is not None:
            dist.all_reduce(de, group=pg)
        de = de.to(e_info.dtype)
    return de, (dc.to(c_info.dtype) if need_c else None), (db.to(bias_info.dtype) if need_b else None)
