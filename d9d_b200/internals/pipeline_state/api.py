from __future__ import annotations

import abc
from typing import Any


class PipelineState(abc.ABC):
    """Dict-like state; whether a key is stored globally or per shard is hidden behind the view."""

    @abc.abstractmethod
    def __setitem__(self, key: str, value: Any) -> None: ...

    @abc.abstractmethod
    def __getitem__(self, item: str) -> Any: ...

    @abc.abstractmethod
    def __contains__(self, item: str) -> bool: ...
