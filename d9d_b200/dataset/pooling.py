from __future__ import annotations

import enum

import torch


class TokenPoolingType(enum.StrEnum):
    first = "first"  # first real token (e.g. [CLS])
    last = "last"  # last real token (decoder embeddings)
    all = "all"  # every real token (mean pooling)


def token_pooling_mask_from_attention_mask(attention_mask: torch.Tensor, pooling_type: TokenPoolingType) -> torch.Tensor:
    """``[B, S]`` 0/1 attention mask -> 0/1 mask of the tokens to pool.

    Works for right- *and* left-padded batches: ``first`` / ``last`` are located from the mask itself (the reference,
    ``d9d/dataset/pooling.py``, assumes right padding: position 0 and ``sum - 1``).  A row without any real token
    selects nothing.
    """
    if pooling_type == TokenPoolingType.all:
        return attention_mask
    if pooling_type not in (TokenPoolingType.first, TokenPoolingType.last):
        raise ValueError(f"Unknown pooling type: {pooling_type}")
    real = attention_mask != 0
    seq = attention_mask.size(1)
    if pooling_type == TokenPoolingType.first:
        position = real.to(torch.int8).argmax(dim=1)
    else:
        position = seq - 1 - real.flip(1).to(torch.int8).argmax(dim=1)
    mask = torch.zeros_like(attention_mask, dtype=torch.long)
    mask.scatter_(1, position[:, None], 1)
    return mask * real.any(dim=1, keepdim=True).long()
