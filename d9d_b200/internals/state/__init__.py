"""State-dict helpers for objects that only live on global rank 0 (reference ``internals/state/main_process.py``)."""

from __future__ import annotations

from typing import Any

from torch.distributed.checkpoint.stateful import Stateful

from d9d_b200.core.dist_context import DistributedContext


def state_dict_main_process(dist_context: DistributedContext, obj: Stateful) -> dict[str, Any]:
    return {"main_process": obj.state_dict()} if dist_context.is_main_process else {}


def load_state_dict_main_process(dist_context: DistributedContext, obj: Stateful, state_dict: dict[str, Any]) -> None:
    if dist_context.is_main_process:
        obj.load_state_dict(state_dict["main_process"])


__all__ = ["load_state_dict_main_process", "state_dict_main_process"]
