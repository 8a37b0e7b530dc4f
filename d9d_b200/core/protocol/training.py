from __future__ import annotations

from typing import Any, Protocol, runtime_checkable


@runtime_checkable
class OptimizerProtocol(Protocol):
    """Anything with the ``torch.optim.Optimizer`` stepping + Stateful surface."""

    def step(self) -> None: ...

    def zero_grad(self) -> None: ...

    def state_dict(self) -> dict[str, Any]: ...

    def load_state_dict(self, state_dict: dict[str, Any]) -> None: ...


@runtime_checkable
class LRSchedulerProtocol(Protocol):
    """Anything with the LR-scheduler stepping + Stateful surface."""

    def step(self) -> None: ...

    def state_dict(self) -> dict[str, Any]: ...

    def load_state_dict(self, state_dict: dict[str, Any]) -> None: ...
