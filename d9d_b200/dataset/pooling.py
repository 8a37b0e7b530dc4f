from __future__ import annotations

import enum

import torch


class TokenPoolingType(enum.StrEnum):
    first = "first"  # e.g. [CLS]
    last = "last"  # last non-padding token (decoder embeddings)
    all = "all"  # every non-padding token (mean pooling)


def token_pooling_mask_from_attention_mask(attention_mask: torch.Tensor, pooling_type: TokenPoolingType) -> torch.Tensor:
    """``[B, S]`` 0/1 mask selecting the tokens to pool for the given strategy."""
    if pooling_type == TokenPoolingType.all:
        return attention_mask
    mask = torch.zeros_like(attention_mask, dtype=torch.long)
    if pooling_type == TokenPoolingType.first:
        mask[:, 0] = 1
        return mask
    if pooling_type == TokenPoolingType.last:
        last = attention_mask.sum(dim=1) - 1
        mask[torch.arange(attention_mask.size(0), device=attention_mask.device), last] = 1
        return mask
    raise ValueError(f"Unknown pooling type: {pooling_type}")
