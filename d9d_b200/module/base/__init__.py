"""Structural protocols for modules (reference ``d9d/module/base/late_init.py``)."""

from typing import Protocol, runtime_checkable


@runtime_checkable
class ModuleLateInit(Protocol):
    """Modules built on the meta device and materialised later implement ``reset_parameters``."""

    def reset_parameters(self) -> None: ...


__all__ = ["ModuleLateInit"]
