"""Aim backend (reference ``d9d/tracker/provider/aim/tracker.py``).

``aim`` is an optional dependency and is imported when a run is opened, so configs mentioning this tracker can be
parsed - and the tracker constructed, checkpointed and restored - on machines without it.
"""

from __future__ import annotations

import importlib
from collections.abc import Generator
from contextlib import contextmanager
from typing import Any, Self, TypedDict

import torch

from d9d_b200.tracker.base import BaseTracker, BaseTrackerRun, RunConfig

from .config import AimConfig


class AimState(TypedDict):
    """Checkpointed tracker state: the hash of the Aim run to continue after a restart."""

    restart_hash: str | None


def _aim() -> Any:
    try:
        return importlib.import_module("aim")
    except ImportError as exc:
        raise ImportError("The Aim tracker needs the optional `aim` package") from exc


class AimRun(BaseTrackerRun):
    """Forwards scalars / histograms to an ``aim.Run``, stamping them with the current step and tag context."""

    def __init__(self, run: Any):
        self._sink = run
        self._step = 0
        self._tags: dict[str, str] = {}

    def set_step(self, step: int) -> None:
        self._step = step

    def set_context(self, context: dict[str, str]) -> None:
        self._tags = dict(context)

    def _emit(self, name: str, value: Any, extra_tags: dict[str, str] | None) -> None:
        tags = self._tags if not extra_tags else self._tags | extra_tags
        self._sink.track(value, name=name, step=self._step, context=tags)

    def scalar(self, name: str, value: float, context: dict[str, str] | None = None) -> None:
        self._emit(name, value, context)

    def bins(self, name: str, values: torch.Tensor, context: dict[str, str] | None = None) -> None:
        counts = values.detach().cpu().numpy()
        self._emit(name, _aim().Distribution(hist=counts, bin_range=(0, counts.shape[0])), context)


class AimTracker(BaseTracker[AimConfig]):
    """Opens Aim runs; the run hash travels in the job checkpoint so that a restarted job appends to the same run."""

    def __init__(self, config: AimConfig):
        self._config = config
        self._state = AimState(restart_hash=None)

    @classmethod
    def from_config(cls, config: AimConfig) -> Self:
        return cls(config)

    def state_dict(self) -> dict[str, Any]:
        return dict(self._state)

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self._state = AimState(restart_hash=state_dict["restart_hash"])

    @contextmanager
    def open(self, properties: RunConfig) -> Generator[BaseTrackerRun, None, None]:
        cfg = self._config
        run = _aim().Run(run_hash=self._state["restart_hash"], repo=cfg.repo, log_system_params=cfg.log_system_params,
                         capture_terminal_logs=cfg.capture_terminal_logs, system_tracking_interval=cfg.system_tracking_interval)
        try:
            run.name, run.description = properties.name, properties.description
            run["hparams"] = properties.hparams
            self._state["restart_hash"] = run.hash
            yield AimRun(run)
        finally:
            run.close()
