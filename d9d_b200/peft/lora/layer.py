from __future__ import annotations

import torch
from torch import nn

from d9d_b200.module.block.linear import Linear
from d9d_b200.module.block.moe import GroupedLinear

from .config import LoRAParameters


class LoRALinear(nn.Module):
    """``base(x) + alpha/r * B(A(dropout(x)))``; A random, B zero at init (reference ``peft/lora/layer.py:9-83``)."""

    def __init__(self, base_layer: nn.Linear, params: LoRAParameters):
        super().__init__()
        if base_layer.bias is not None:
            raise ValueError("LoRA is unsupported with biased linear layers")
        kw = dict(bias=False, device=base_layer.weight.device, dtype=base_layer.weight.dtype)
        # the rank is tiny (8..64): the adapter GEMMs stay on plain nn.Linear when r is below the native GEMM's granularity
        cls = Linear if params.r % 8 == 0 else nn.Linear
        self.lora_A = cls(base_layer.in_features, params.r, **kw)
        self.lora_B = cls(params.r, base_layer.out_features, **kw)
        self.base = base_layer
        self.dropout = nn.Dropout(params.dropout)
        self._scale = params.alpha / params.r
        self._reset_adapters()  # wrapping an initialised layer must not touch its weights

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.base(x) + self._scale * self.lora_B(self.lora_A(self.dropout(x)))

    @torch.no_grad()
    def merge_with_base_(self) -> nn.Linear:
        self.base.weight.data += (self.lora_B.weight.data.float() @ self.lora_A.weight.data.float()).to(self.base.weight.dtype) * self._scale
        return self.base

    def reset_parameters(self) -> None:
        """Late initialisation of the whole wrapped layer: the (frozen) base as well - a model built on the meta device and
        materialised with ``to_empty`` has no valid base weights until they are initialised here or loaded."""
        if self.lora_A.weight.is_meta:
            return
        self.base.reset_parameters()
        self._reset_adapters()

    def _reset_adapters(self) -> None:
        if self.lora_A.weight.is_meta:
            return
        self.lora_A.reset_parameters()
        nn.init.zeros_(self.lora_B.weight)


class LoRAGroupedLinear(nn.Module):
    """LoRA over every expert of a ``GroupedLinear`` (adapter weights ``[E, in, r]`` and ``[E, r, out]``)."""

    def __init__(self, base_layer: GroupedLinear, params: LoRAParameters):
        super().__init__()
        kw = dict(device=base_layer.weight.device, dtype=base_layer.weight.dtype)
        self.lora_A = GroupedLinear(base_layer.n_groups, base_layer.in_features, params.r, **kw)
        self.lora_B = GroupedLinear(base_layer.n_groups, params.r, base_layer.out_features, **kw)
        self.base = base_layer
        self.dropout = nn.Dropout(params.dropout)
        self._scale = params.alpha / params.r
        self._reset_adapters()  # wrapping an initialised layer must not touch its weights

    def forward(self, x: torch.Tensor, x_groups) -> torch.Tensor:
        return self.base(x, x_groups) + self._scale * self.lora_B(self.lora_A(self.dropout(x), x_groups), x_groups)

    @torch.no_grad()
    def merge_with_base_(self) -> GroupedLinear:
        delta = torch.bmm(self.lora_A.weight.data.float(), self.lora_B.weight.data.float()) * self._scale
        self.base.weight.data += delta.to(self.base.weight.dtype)
        return self.base

    def reset_parameters(self) -> None:
        """Late initialisation of the whole wrapped layer: the (frozen) base as well - a model built on the meta device and
        materialised with ``to_empty`` has no valid base weights until they are initialised here or loaded."""
        if self.lora_A.weight.is_meta:
            return
        self.base.reset_parameters()
        self._reset_adapters()

    def _reset_adapters(self) -> None:
        if self.lora_A.weight.is_meta:
            return
        self.lora_A.reset_parameters()
        nn.init.zeros_(self.lora_B.weight)
