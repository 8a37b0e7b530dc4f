from __future__ import annotations

import torch

from d9d_b200.kernel.moe import MoELayout, build_moe_layout, moe_permute, moe_unpermute

from .base import ExpertCommunicationHandler


class NoCommunicationHandler(ExpertCommunicationHandler):
    """All experts are local: build the aligned layout on the device and permute (no collective, no host sync).

    Parity: reference ``moe/communications/naive.py:29-45`` (which does ``bincount().cpu()`` + three Triton passes).
    """

    def __init__(self, num_experts: int):
        self._num_experts = num_experts
        self._layout: MoELayout | None = None

    def dispatch(self, hidden_states: torch.Tensor, topk_ids: torch.Tensor, topk_weights: torch.Tensor):
        layout = build_moe_layout(topk_ids, self._num_experts)
        xp, pp = moe_permute(hidden_states, topk_weights, layout)
        self._layout = layout
        return xp, pp, layout

    def combine(self, hidden_states: torch.Tensor) -> torch.Tensor:
        if self._layout is None:
            raise ValueError("Cannot run combine before running dispatch!")
        layout, self._layout = self._layout, None
        return moe_unpermute(hidden_states, layout)
