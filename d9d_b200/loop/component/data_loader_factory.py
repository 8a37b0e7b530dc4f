from __future__ import annotations

import pickle
from collections.abc import Iterator
from typing import Any, Self

import torch
import torch.utils._pytree as pytree
from torch.utils.data import Dataset
from torchdata.stateful_dataloader import StatefulDataLoader

from d9d_b200.core.dist_context import BATCH_DOMAIN, DistributedContext
from d9d_b200.loop.config import DataLoadingConfig
from d9d_b200.loop.control import DatasetProvider, InitializeDatasetContext

from .batch_maths import BatchMaths


def _to_device(tree: Any, device: torch.device) -> Any:
    non_blocking = device.type == "cuda"  # pinned host memory -> async H2D copy on the current stream
    return pytree.tree_map(lambda x: x.to(device, non_blocking=non_blocking) if isinstance(x, torch.Tensor) else x, tree)


class IteratorBatchGroup(Iterator):
    """Yields *groups*: each item is an iterator over ``batch_group_size`` consecutive batches (the gradient
    accumulation rounds of one optimizer step), moved to the device as they are consumed."""

    def __init__(self, base: Iterator, device: torch.device, batch_group_size: int):
        self._base = base
        self._device = device
        self._group = batch_group_size
        self._exhausted = False

    def __iter__(self) -> Self:
        return self

    def __next__(self) -> Iterator[Any]:
        if self._exhausted:
            raise StopIteration
        try:
            first = next(self._base)
        except StopIteration:
            self._exhausted = True
            raise

        def group() -> Iterator[Any]:
            yield _to_device(first, self._device)
            for _ in range(self._group - 1):
                try:
                    item = next(self._base)
                except StopIteration:
                    self._exhausted = True
                    return
                yield _to_device(item, self._device)

        return group()


class StatefulDataLoaderDataParallelAware(StatefulDataLoader):
    """``StatefulDataLoader`` whose checkpoint state is keyed by data-parallel rank (``dp_{rank}``) and whose iterator
    yields device-resident batch groups."""

    def __init__(self, dataset: Dataset, dp_rank: int, device: torch.device | str, group_size: int, **kwargs: Any):
        super().__init__(dataset, **kwargs)
        self._dp_rank = dp_rank
        self._target_device = torch.device(device)
        self._group_size = group_size

    def state_dict(self) -> dict[str, Any]:
        # the loader state is an opaque nested structure whose shape changes between a fresh and a running loader;
        # checkpointing it as one pickled blob keeps the DCP key set stable
        return {f"dp_{self._dp_rank}": pickle.dumps(super().state_dict())}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        blob = state_dict[f"dp_{self._dp_rank}"]
        super().load_state_dict(pickle.loads(blob) if isinstance(blob, (bytes, bytearray)) else blob)  # noqa: S301

    def __iter__(self) -> Iterator:  # type: ignore[override]
        return IteratorBatchGroup(super().__iter__(), device=self._target_device, batch_group_size=self._group_size)


class DataLoaderFactory:
    def __init__(self, dist_context: DistributedContext, provider: DatasetProvider, config_data_loading: DataLoadingConfig,
                 batch_maths: BatchMaths):
        self._ctx = dist_context
        self._provider = provider
        self._config = config_data_loading
        self._maths = batch_maths

    def _build(self, batch_size: int, group_size: int, drop_last: bool) -> StatefulDataLoaderDataParallelAware:
        result = self._provider(InitializeDatasetContext(dist_context=self._ctx, batch_maths=self._maths))
        dp_rank = self._ctx.mesh_for(BATCH_DOMAIN)["dp"].get_local_rank() if self._ctx.mesh_params.is_distributed else 0
        on_cuda = self._ctx.current_device.type == "cuda"
        return StatefulDataLoaderDataParallelAware(
            result.dataset,
            dp_rank=dp_rank,
            device=self._ctx.current_device,
            group_size=group_size,
            collate_fn=result.collator,
            batch_size=batch_size,
            drop_last=drop_last,
            num_workers=self._config.num_workers,
            persistent_workers=self._config.persistent_workers and self._config.num_workers > 0,
            pin_memory=self._config.pin_memory and on_cuda,
        )

    def build_dataloader_for_train_job(self) -> StatefulDataLoaderDataParallelAware:
        return self._build(self._maths.data_loader_batch_size, self._maths.num_microbatches_gradient_accumulation, drop_last=True)

    def build_dataloader_for_infer_job(self) -> StatefulDataLoaderDataParallelAware:
        return self._build(self._maths.data_loader_batch_size, 1, drop_last=False)
