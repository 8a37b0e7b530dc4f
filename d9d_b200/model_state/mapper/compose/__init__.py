"""Composite mappers (reference ``d9d/model_state/mapper/compose``)."""

from __future__ import annotations

from collections.abc import Callable, Sequence

import torch

from d9d_b200.model_state.mapper.abc import ModelStateMapper, StateGroup
from d9d_b200.model_state.mapper.leaf import ModelStateMapperIdentity


def filter_empty_mappers(mappers: Sequence[ModelStateMapper]) -> list[ModelStateMapper]:
    """Drop mappers whose every group has neither inputs nor outputs."""
    return [m for m in mappers if any(g.inputs or g.outputs for g in m.state_dependency_groups())]


class ModelStateMapperParallel(ModelStateMapper):
    """Independent mappers side by side.  Input keys and output keys must be globally unique; ``apply`` is routed by
    the exact set of input keys."""

    def __init__(self, mappers: Sequence[ModelStateMapper]):
        route: dict[frozenset[str], ModelStateMapper] = {}
        groups: set[StateGroup] = set()
        taken_in: set[str] = set()
        taken_out: set[str] = set()
        for mapper in filter_empty_mappers(mappers):
            for g in mapper.state_dependency_groups():
                if taken_in & g.inputs:
                    raise ValueError(f"Found a colliding input group: {g.inputs}")
                if taken_out & g.outputs:
                    raise ValueError(f"Found colliding output keys: {g.outputs}")
                taken_in |= g.inputs
                taken_out |= g.outputs
                groups.add(g)
                route[g.inputs] = mapper
        self._groups = frozenset(groups)
        self._route = route

    def state_dependency_groups(self) -> frozenset[StateGroup]:
        return self._groups

    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        mapper = self._route.get(frozenset(group))
        if mapper is None:
            raise ValueError("Tried to run a parallel mapper with undefined group. Perhaps you sent groups that are not isolated?")
        return mapper.apply(group)


class ModelStateMapperPrefixScope(ModelStateMapper):
    """Run a child mapper under independent source / target key prefixes."""

    def __init__(self, mapper: ModelStateMapper, source_prefix: str = "", target_prefix: str = "") -> None:
        self._mapper, self._src, self._dst = mapper, source_prefix, target_prefix
        self._groups = frozenset(
            StateGroup(inputs=frozenset(source_prefix + k for k in g.inputs),
                       outputs=frozenset(target_prefix + k for k in g.outputs))
            for g in mapper.state_dependency_groups()
        )

    def state_dependency_groups(self) -> frozenset[StateGroup]:
        return self._groups

    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        inner = self._mapper.apply({k.removeprefix(self._src): v for k, v in group.items()})
        return {self._dst + k: v for k, v in inner.items()}


class ModelStateMapperShard(ModelStateMapper):
    """Keep every ``total_shards``-th dependency group (deterministic order) — splits loading work over processes."""

    def __init__(self, sub_mapper: ModelStateMapper, total_shards: int, current_shard: int):
        ordered = sorted(sub_mapper.state_dependency_groups(), key=lambda g: sorted(g.inputs))
        self._groups = frozenset(g for i, g in enumerate(ordered) if i % total_shards == current_shard)
        self._sub = sub_mapper

    def state_dependency_groups(self) -> frozenset[StateGroup]:
        return self._groups

    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        return self._sub.apply(group)


class ModelStateMapperSelectGroups(ModelStateMapper):
    """Keep the dependency groups accepted by ``predicate`` - e.g. the part of a whole-model mapper that concerns the
    parameters of one pipeline stage."""

    def __init__(self, sub_mapper: ModelStateMapper, predicate: Callable[[StateGroup], bool]):
        self._groups = frozenset(g for g in sub_mapper.state_dependency_groups() if predicate(g))
        self._sub = sub_mapper

    def state_dependency_groups(self) -> frozenset[StateGroup]:
        return self._groups

    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        return self._sub.apply(group)


def _keys(groups: frozenset[StateGroup], attr: str) -> set[str]:
    out: set[str] = set()
    for g in groups:
        out |= getattr(g, attr)
    return out


class ModelStateMapperSequential(ModelStateMapper):
    """Mappers chained stage after stage.

    * gap filling: keys a later stage needs (or an earlier stage produced) that an intermediate stage does not
      touch are carried through by injected identities;
    * the *net* dependency groups are the connected components of the stage-to-stage group graph: inputs of the
      first stage on one side, outputs of the last stage on the other.
    """

    def __init__(self, mappers: list[ModelStateMapper]):
        stages = filter_empty_mappers(mappers)
        if not stages:
            raise ValueError("Mappers list cannot be empty.")
        stages = self._fill_gaps(stages)
        self._stages = stages
        self._groups = self._net_groups(stages)

    @staticmethod
    def _fill_gaps(stages: list[ModelStateMapper]) -> list[ModelStateMapper]:
        stages = list(stages)
        for i in range(len(stages) - 1, 0, -1):  # what later stages need must flow through earlier ones
            missing = _keys(stages[i].state_dependency_groups(), "inputs") - _keys(stages[i - 1].state_dependency_groups(), "outputs")
            if missing:
                stages[i - 1] = ModelStateMapperParallel([stages[i - 1], *(ModelStateMapperIdentity(k) for k in sorted(missing))])
        for i in range(len(stages) - 1):  # what earlier stages produced must survive until the end
            extra = _keys(stages[i].state_dependency_groups(), "outputs") - _keys(stages[i + 1].state_dependency_groups(), "inputs")
            if extra:
                stages[i + 1] = ModelStateMapperParallel([stages[i + 1], *(ModelStateMapperIdentity(k) for k in sorted(extra))])
        return stages

    @staticmethod
    def _net_groups(stages: list[ModelStateMapper]) -> frozenset[StateGroup]:
        nodes: list[tuple[int, StateGroup]] = [(s, g) for s, m in enumerate(stages) for g in m.state_dependency_groups()]
        parent = list(range(len(nodes)))

        def find(x: int) -> int:
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x

        def union(a: int, b: int) -> None:
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[rb] = ra

        producer: dict[tuple[int, str], int] = {}
        for idx, (s, g) in enumerate(nodes):
            for key in g.outputs:
                producer[(s, key)] = idx
        first_in: dict[str, int] = {}
        last_out: dict[str, int] = {}
        last = len(stages) - 1
        for idx, (s, g) in enumerate(nodes):
            if s > 0:
                for key in g.inputs:
                    src = producer.get((s - 1, key))
                    if src is not None:
                        union(idx, src)
            if s == 0:  # groups sharing an input / output key belong together as well
                for key in g.inputs:
                    if key in first_in:
                        union(idx, first_in[key])
                    first_in[key] = idx
            if s == last:
                for key in g.outputs:
                    if key in last_out:
                        union(idx, last_out[key])
                    last_out[key] = idx
        comps: dict[int, tuple[set[str], set[str]]] = {}
        for idx, (s, g) in enumerate(nodes):
            ins, outs = comps.setdefault(find(idx), (set(), set()))
            if s == 0:
                ins |= g.inputs
            if s == last:
                outs |= g.outputs
        return frozenset(StateGroup(inputs=frozenset(i), outputs=frozenset(o)) for i, o in comps.values() if i or o)

    def state_dependency_groups(self) -> frozenset[StateGroup]:
        return self._groups

    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        state = group
        for stage in self._stages:
            produced: dict[str, torch.Tensor] = {}
            for g in stage.state_dependency_groups():
                if g.inputs <= state.keys():
                    produced.update(stage.apply({k: state[k] for k in g.inputs}))
            state = produced
        return state


__all__ = [
    "ModelStateMapperParallel",
    "ModelStateMapperPrefixScope",
    "ModelStateMapperSelectGroups",
    "ModelStateMapperSequential",
    "ModelStateMapperShard",
    "filter_empty_mappers",
]
