from __future__ import annotations

import torch
import torch.nn.functional as F

from .utils import handle_reduction_none  # noqa: F401


def torch_compile_linear_cross_entropy(e, c, targets, bias=None, ignore_index=-100, softcap=None, reduction="mean", shift=0,
                                       vocab_parallel_options=None, return_lse=False):
    """Plain (uncompiled) PyTorch fallback with the real function's signature."""
    if vocab_parallel_options is not None:
        raise NotImplementedError("vocab parallelism is not supported by the stand-in torch path")
    shift = int(shift)
    if shift:
        e, targets = e[..., :-shift, :], targets[..., shift:]
    logits = (e @ c.t()).float()
    if bias is not None:
        logits = logits + bias.float()
    if softcap is not None:
        logits = torch.tanh(logits / softcap) * softcap
    loss = F.cross_entropy(logits.flatten(0, -2), targets.flatten(), ignore_index=ignore_index, reduction=reduction)
    if reduction == "none":
        loss = loss.view(targets.shape)
    lse = torch.logsumexp(logits, dim=-1) if return_lse else None
    return loss, lse
