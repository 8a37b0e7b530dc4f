from __future__ import annotations

import abc

import torch

from d9d_b200.kernel.moe import MoELayout


class ExpertCommunicationHandler(abc.ABC):
    """Moves tokens to the experts that own them and back (reference ``moe/communications/base.py:10-52``).

    ``dispatch`` returns ``(rows sorted by local expert, matching routing probabilities, grouping)`` where the
    grouping is a device-side :class:`MoELayout` (the reference returns a CPU ``tokens_per_expert`` tensor, which
    costs a host sync per layer).  Handlers are stateful between ``dispatch`` and ``combine``.
    """

    @abc.abstractmethod
    def dispatch(self, hidden_states: torch.Tensor, topk_ids: torch.Tensor, topk_weights: torch.Tensor
                 ) -> tuple[torch.Tensor, torch.Tensor, MoELayout | torch.Tensor]: ...

    @abc.abstractmethod
    def combine(self, hidden_states: torch.Tensor) -> torch.Tensor: ...

    def expert_buffers(self) -> tuple[torch.Tensor | None, torch.Tensor | None]:
        """Between ``dispatch`` and ``combine``: optional ``[>= capacity, hidden]`` buffers the expert block may write its
        output / its input gradient into directly (memory the handler would otherwise have to copy them to)."""
        return None, None
