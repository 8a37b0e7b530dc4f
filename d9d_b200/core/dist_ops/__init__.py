"""Convenience collectives that allocate their outputs (reference ``d9d/core/dist_ops``).

Differences from the reference: shapes of variadic tensors are exchanged in a *single* fixed-size all-gather
(``[ndim, d0..d7]`` per rank) instead of two dependent rounds, and everything works on gloo/CPU as well as NCCL.
"""

from .object import all_gather_object, gather_object
from .tensor import all_gather, all_gather_variadic_shape, gather, gather_variadic_shape

__all__ = [
    "all_gather",
    "all_gather_object",
    "all_gather_variadic_shape",
    "gather",
    "gather_object",
    "gather_variadic_shape",
]
