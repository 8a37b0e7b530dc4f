from __future__ import annotations

import abc
import dataclasses

import torch


@dataclasses.dataclass(frozen=True)
class StateGroup:
    """Atomic dependency: the ``outputs`` keys can be produced once all ``inputs`` keys are available."""

    inputs: frozenset[str]
    outputs: frozenset[str]


class ModelStateMapper(abc.ABC):
    """Transformation of (a subset of) a state dict.

    ``state_dependency_groups`` is pure topology (no tensors touched) so readers/writers can plan IO, validate
    chains and shard work before allocating memory; ``apply`` runs the tensor ops for ONE group.
    """

    @abc.abstractmethod
    def state_dependency_groups(self) -> frozenset[StateGroup]: ...

    @abc.abstractmethod
    def apply(self, group: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]: ...

    # convenience used by readers / tests
    def all_inputs(self) -> frozenset[str]:
        out: set[str] = set()
        for g in self.state_dependency_groups():
            out |= g.inputs
        return frozenset(out)

    def all_outputs(self) -> frozenset[str]:
        out: set[str] = set()
        for g in self.state_dependency_groups():
            out |= g.outputs
        return frozenset(out)
