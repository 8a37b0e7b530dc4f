from __future__ import annotations

from collections.abc import Generator
from contextlib import contextmanager
from typing import Any, Self

import torch
from aim import Distribution, Run  # optional dependency: importing this module without aim raises ImportError

from d9d_b200.tracker.base import BaseTracker, BaseTrackerRun, RunConfig

from .config import AimConfig


class AimRun(BaseTrackerRun):
    def __init__(self, run: Run):
        self._run = run
        self._step = 0
        self._context: dict[str, str] = {}

    def set_step(self, step: int) -> None:
        self._step = step

    def set_context(self, context: dict[str, str]) -> None:
        self._context = context

    def _ctx(self, context: dict[str, str] | None) -> dict[str, str]:
        return self._context if context is None else {**self._context, **context}

    def scalar(self, name: str, value: float, context: dict[str, str] | None = None) -> None:
        self._run.track(name=name, value=value, context=self._ctx(context), step=self._step)

    def bins(self, name: str, values: torch.Tensor, context: dict[str, str] | None = None) -> None:
        self._run.track(name=name, value=Distribution(hist=values.numpy(), bin_range=(0, values.shape[0])),
                        context=self._ctx(context), step=self._step)


class AimTracker(BaseTracker[AimConfig]):
    """Aim tracker that remembers the run hash in its state dict so a restarted job resumes the same run."""

    def __init__(self, config: AimConfig):
        self._config = config
        self._restart_hash: str | None = None

    def state_dict(self) -> dict[str, Any]:
        return {"restart_hash": self._restart_hash}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self._restart_hash = state_dict["restart_hash"]

    @contextmanager
    def open(self, properties: RunConfig) -> Generator[BaseTrackerRun, None, None]:
        run = Run(run_hash=self._restart_hash, repo=self._config.repo, log_system_params=self._config.log_system_params,
                  capture_terminal_logs=self._config.capture_terminal_logs,
                  system_tracking_interval=self._config.system_tracking_interval)
        run.name = properties.name
        run.description = properties.description
        run["hparams"] = properties.hparams
        self._restart_hash = run.hash
        try:
            yield AimRun(run)
        finally:
            run.close()

    @classmethod
    def from_config(cls, config: AimConfig) -> Self:
        return cls(config)
