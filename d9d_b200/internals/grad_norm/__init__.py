"""Distributed gradient-norm computation and clipping (reference ``d9d/internals/grad_norm``)."""

from .group import GradNormGroup, ParametersForNorm, group_parameters_for_norm
from .norm import clip_grad_norm_distributed_

__all__ = ["GradNormGroup", "ParametersForNorm", "clip_grad_norm_distributed_", "group_parameters_for_norm"]
