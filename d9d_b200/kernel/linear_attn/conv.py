from __future__ import annotations

import torch
import torch.nn.functional as F


def causal_conv1d_silu(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """Depthwise causal convolution followed by SiLU.

    ``x [B, S, C]``, ``weight [C, K]`` (tap ``K-1`` multiplies the current position) ->
    ``y[b, s, c] = silu(sum_j weight[c, j] * x[b, s - (K-1) + j, c])`` with zeros left of the sequence.
    """
    k = weight.shape[1]
    xt = F.pad(x.transpose(1, 2), (k - 1, 0))
    y = F.conv1d(xt, weight.unsqueeze(1).to(x.dtype), groups=weight.shape[0])
    return F.silu(y).transpose(1, 2)
