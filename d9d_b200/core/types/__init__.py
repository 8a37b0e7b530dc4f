"""Common type aliases (reference ``d9d/core/types``)."""

from collections.abc import Callable, Sequence
from typing import TypeAlias, TypeVar

import torch

TLeaf = TypeVar("TLeaf")

PyTree: TypeAlias = TLeaf | list["PyTree[TLeaf]"] | dict[str, "PyTree[TLeaf]"] | tuple["PyTree[TLeaf]", ...]
TensorTree: TypeAlias = PyTree[torch.Tensor]
ScalarTree: TypeAlias = PyTree[str | float | int | bool]

TDataTree = TypeVar("TDataTree", bound=PyTree)
CollateFn: TypeAlias = Callable[[Sequence[TDataTree]], TDataTree]

__all__ = ["CollateFn", "PyTree", "ScalarTree", "TensorTree"]
