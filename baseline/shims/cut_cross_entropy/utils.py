from __future__ import annotations

from dataclasses import dataclass

import torch


class CCEWarning(UserWarning):
    pass


@dataclass
class TensorInfo:
    dtype: torch.dtype
    requires_grad: bool


def is_torch_greater_or_equal_2_5() -> bool:
    return True


def is_triton_3_2() -> bool:
    return False


def to_full_tensor(t):
    if t is not None and isinstance(t, torch.distributed.tensor.DTensor):
        return t.full_tensor()
    return t


def maybe_type_as(t, other):
    if t is None or t.dtype == other.dtype:
        return t
    return t.type_as(other)


def _handle_eps(filter_eps, dtype: torch.dtype):
    if filter_eps is None:
        return None
    if isinstance(filter_eps, float):
        return filter_eps
    if filter_eps == "auto":
        return torch.finfo(dtype).eps / 32
    raise RuntimeError(f"Unknown eps {filter_eps}")


def _build_flat_valids(targets: torch.Tensor, ignore_index: int, shift: int) -> torch.Tensor | None:
    """Indices (int32, into the flattened un-shifted token axis minus ``shift``) of the rows that carry a loss."""
    if shift != 0:
        targets = targets[..., shift:]
        flat = targets.contiguous().flatten()
        seq = targets.size(-1)
        idx = (flat != ignore_index).nonzero().squeeze(1)
        # position in the un-shifted flattened [*, S] embedding matrix of the row predicting this target
        full = idx // seq * (seq + shift) + idx % seq
        return full.to(torch.int32)
    flat = targets.flatten()
    idx = (flat != ignore_index).nonzero().squeeze(1)
    if idx.numel() == flat.numel():
        return None
    return idx.to(torch.int32)


def handle_reduction_none(batch_shape: torch.Size, valids: torch.Tensor | None, shift: int, value: torch.Tensor) -> torch.Tensor:
    if valids is None:
        return value.view(batch_shape)
    full = value.new_zeros(batch_shape.numel())
    full[(valids + shift).long()] = value
    return full.view(batch_shape)
