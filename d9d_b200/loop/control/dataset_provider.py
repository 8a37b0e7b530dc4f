from __future__ import annotations

import dataclasses
from typing import TYPE_CHECKING, Protocol, runtime_checkable

from torch.utils.data import Dataset

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.core.types import CollateFn

if TYPE_CHECKING:
    from d9d_b200.loop.component import BatchMaths


@dataclasses.dataclass(kw_only=True)
class InitializeDatasetContext:
    dist_context: DistributedContext
    batch_maths: "BatchMaths"


@dataclasses.dataclass(kw_only=True)
class InitializeDatasetResult:
    dataset: Dataset
    collator: CollateFn


@runtime_checkable
class DatasetProvider(Protocol):
    """Builds the (already data-parallel sharded — e.g. with ``shard_dataset_data_parallel``) dataset and its collator."""

    def __call__(self, context: InitializeDatasetContext) -> InitializeDatasetResult: ...
