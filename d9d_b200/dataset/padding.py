from __future__ import annotations

import enum
from collections.abc import Sequence

import torch


class PaddingSide1D(enum.StrEnum):
    left = "left"
    right = "right"


def pad_stack_1d(items: Sequence[torch.Tensor], pad_value: int, padding_side: PaddingSide1D = PaddingSide1D.right,
                 pad_to_multiple_of: int | None = None) -> torch.Tensor:
    """Stack 1-D tensors into ``[batch, max_len]`` (``max_len`` optionally rounded up), padding with ``pad_value``.

    Writes every row once into a pre-filled output instead of padding each item separately.
    """
    if not items:
        raise ValueError("Cannot stack 0 items")
    if pad_to_multiple_of is not None and pad_to_multiple_of <= 0:
        raise ValueError("pad_to_multiple_of should be > 0")
    if padding_side not in (PaddingSide1D.left, PaddingSide1D.right):
        raise ValueError("Unknown padding side")
    width = max(x.shape[0] for x in items)
    if pad_to_multiple_of is not None:
        width = -(-width // pad_to_multiple_of) * pad_to_multiple_of
    out = items[0].new_full((len(items), width), pad_value)
    for row, x in enumerate(items):
        n = x.shape[0]
        if padding_side == PaddingSide1D.right:
            out[row, :n] = x
        else:
            out[row, width - n :] = x
    return out
