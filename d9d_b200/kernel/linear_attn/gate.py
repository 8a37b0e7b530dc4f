from __future__ import annotations

import torch
import torch.nn.functional as F


def mamba_decay_gate(x: torch.Tensor, a_log: torch.Tensor, dt_bias: torch.Tensor) -> torch.Tensor:
    """``g = -exp(A_log) * softplus(x + dt_bias)`` per head, computed in fp32 (log-space decay, always <= 0).

    ``x [..., H]``; ``a_log``, ``dt_bias`` ``[H]``.
    """
    return -(a_log.float().exp()) * F.softplus(x.float() + dt_bias.float())
