from __future__ import annotations

import json
import time
import uuid
from collections.abc import Generator
from contextlib import contextmanager
from pathlib import Path
from typing import Any, Literal, Self

import torch
from pydantic import BaseModel

from d9d_b200.tracker.base import BaseTracker, BaseTrackerRun, RunConfig


class JsonlTrackerConfig(BaseModel):
    """Append-only JSON-lines tracker: works offline, greppable, resumable (same file after a restart)."""

    provider: Literal["jsonl"] = "jsonl"
    directory: str


class JsonlRun(BaseTrackerRun):
    def __init__(self, path: Path):
        self._fh = path.open("a", encoding="utf-8")
        self._step = 0
        self._context: dict[str, str] = {}

    def set_step(self, step: int) -> None:
        self._step = step

    def set_context(self, context: dict[str, str]) -> None:
        self._context = dict(context)

    def _emit(self, record: dict[str, Any], context: dict[str, str] | None) -> None:
        record.update(step=self._step, time=time.time(), context={**self._context, **(context or {})})
        self._fh.write(json.dumps(record) + "\n")
        self._fh.flush()

    def scalar(self, name: str, value: float, context: dict[str, str] | None = None) -> None:
        self._emit({"kind": "scalar", "name": name, "value": float(value)}, context)

    def bins(self, name: str, values: torch.Tensor, context: dict[str, str] | None = None) -> None:
        self._emit({"kind": "bins", "name": name, "values": values.detach().cpu().tolist()}, context)

    def close(self) -> None:
        self._fh.close()


class JsonlTracker(BaseTracker[JsonlTrackerConfig]):
    def __init__(self, config: JsonlTrackerConfig):
        self._config = config
        self._run_id: str | None = None

    @contextmanager
    def open(self, properties: RunConfig) -> Generator[BaseTrackerRun, None, None]:
        directory = Path(self._config.directory)
        directory.mkdir(parents=True, exist_ok=True)
        fresh = self._run_id is None
        if fresh:
            self._run_id = uuid.uuid4().hex[:12]
        run = JsonlRun(directory / f"{properties.name}-{self._run_id}.jsonl")
        if fresh:
            run._emit({"kind": "run", "name": properties.name, "description": properties.description,  # noqa: SLF001
                       "hparams": properties.hparams}, None)
        try:
            yield run
        finally:
            run.close()

    @classmethod
    def from_config(cls, config: JsonlTrackerConfig) -> Self:
        return cls(config)

    def state_dict(self) -> dict[str, Any]:
        return {"run_id": self._run_id}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self._run_id = state_dict.get("run_id")
