"""Leaf mappers (reference ``d9d/model_state/mapper/leaf``)."""

from __future__ import annotations

from collections.abc import Callable, Sequence

import torch
from torch.distributed import DeviceMesh
from torch.distributed.tensor import DTensor, Placement, distribute_tensor

from d9d_b200.model_state.mapper.abc import ModelStateMapper, StateGroup


def _one_to_one(src: str, dst: str) -> frozenset[StateGroup]:
    return frozenset((StateGroup(inputs=frozenset((src,)), outputs=frozenset((dst,))),))


class _InPlaceTensorMapper(ModelStateMapper):
    """One key in, same key out, tensor transformed by ``fn``."""

    def __init__(self, name: str, fn: Callable[[torch.Tensor], torch.Tensor]):
        self._name = name
        self._fn = fn
        self._groups = _one_to_one(name, name)

    def state_dependency_groups(self) -> frozenset[StateGroup]:
        return self._groups

    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        return {self._name: self._fn(group[self._name])}


class ModelStateMapperIdentity(_InPlaceTensorMapper):
    """Pass a tensor through unchanged."""

    def __init__(self, name: str):
        super().__init__(name, lambda t: t)


class ModelStateMapperTranspose(_InPlaceTensorMapper):
    def __init__(self, name: str, dims: tuple[int, int]):
        super().__init__(name, lambda t: t.transpose(*dims).contiguous())


class ModelStateMapperSqueeze(_InPlaceTensorMapper):
    def __init__(self, name: str, dim: int | None = None):
        super().__init__(name, (lambda t: t.squeeze()) if dim is None else (lambda t: t.squeeze(dim)))


class ModelStateMapperUnsqueeze(_InPlaceTensorMapper):
    def __init__(self, name: str, dim: int):
        super().__init__(name, lambda t: t.unsqueeze(dim))


class ModelStateMapperRename(ModelStateMapper):
    def __init__(self, name_from: str, name_to: str):
        self._src, self._dst = name_from, name_to
        self._groups = _one_to_one(name_from, name_to)

    def state_dependency_groups(self) -> frozenset[StateGroup]:
        return self._groups

    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        return {self._dst: group[self._src]}


class ModelStateMapperSelectChildModules(ModelStateMapper):
    """Hoist ``{parent}.{name}`` keys to ``{name}`` (one independent group per key)."""

    def __init__(self, base_names: list[str], parent_name: str):
        self._prefix = f"{parent_name}."
        self._groups = frozenset(
            StateGroup(inputs=frozenset((self._prefix + n,)), outputs=frozenset((n,))) for n in base_names
        )

    def state_dependency_groups(self) -> frozenset[StateGroup]:
        return self._groups

    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        return {k[len(self._prefix):]: v for k, v in group.items() if k.startswith(self._prefix)}


class _ManyToOne(ModelStateMapper):
    def __init__(self, source_names: list[str], target_name: str, fn: Callable[[list[torch.Tensor]], torch.Tensor]):
        self._sources, self._target, self._fn = list(source_names), target_name, fn
        self._groups = frozenset((StateGroup(inputs=frozenset(source_names), outputs=frozenset((target_name,))),))

    def state_dependency_groups(self) -> frozenset[StateGroup]:
        return self._groups

    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        return {self._target: self._fn([group[n] for n in self._sources]).contiguous()}


class _OneToMany(ModelStateMapper):
    def __init__(self, source_name: str, target_names: list[str], fn: Callable[[torch.Tensor, int], Sequence[torch.Tensor]]):
        self._source, self._targets, self._fn = source_name, list(target_names), fn
        self._groups = frozenset((StateGroup(inputs=frozenset((source_name,)), outputs=frozenset(target_names)),))

    def state_dependency_groups(self) -> frozenset[StateGroup]:
        return self._groups

    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        pieces = self._fn(group[self._source], len(self._targets))
        return {n: p.contiguous() for n, p in zip(self._targets, pieces, strict=True)}


class ModelStateMapperStackTensors(_ManyToOne):
    def __init__(self, source_names: list[str], target_name: str, dim: int):
        super().__init__(source_names, target_name, lambda ts: torch.stack(ts, dim=dim))


class ModelStateMapperConcatenateTensors(_ManyToOne):
    def __init__(self, source_names: list[str], target_name: str, dim: int):
        super().__init__(source_names, target_name, lambda ts: torch.cat(ts, dim=dim))


class ModelStateMapperUnstackTensors(_OneToMany):
    def __init__(self, source_name: str, target_names: list[str], dim: int):
        super().__init__(source_name, target_names, lambda t, _n: torch.unbind(t, dim=dim))


class ModelStateMapperChunkTensors(_OneToMany):
    def __init__(self, source_name: str, target_names: list[str], dim: int):
        super().__init__(source_name, target_names, lambda t, n: torch.chunk(t, chunks=n, dim=dim))


class ModelStateMapperDistribute(_InPlaceTensorMapper):
    """Local full tensor -> DTensor with the given placements; every rank slices its own copy (no communication)."""

    def __init__(self, name: str, device_mesh: DeviceMesh | None, placements: Sequence[Placement] | None):
        super().__init__(name, lambda t: distribute_tensor(t, device_mesh=device_mesh, placements=placements, src_data_rank=None))


class ModelStateMapperGatherFullTensor(_InPlaceTensorMapper):
    """DTensor -> full local tensor (collective)."""

    def __init__(self, name: str):
        def gather(t: torch.Tensor) -> torch.Tensor:
            if not isinstance(t, DTensor):
                raise ValueError("Cannot gather anything but DTensor")
            return t.full_tensor()

        super().__init__(name, gather)


__all__ = [
    "ModelStateMapperChunkTensors",
    "ModelStateMapperConcatenateTensors",
    "ModelStateMapperDistribute",
    "ModelStateMapperGatherFullTensor",
    "ModelStateMapperIdentity",
    "ModelStateMapperRename",
    "ModelStateMapperSelectChildModules",
    "ModelStateMapperSqueeze",
    "ModelStateMapperStackTensors",
    "ModelStateMapperTranspose",
    "ModelStateMapperUnsqueeze",
    "ModelStateMapperUnstackTensors",
]
