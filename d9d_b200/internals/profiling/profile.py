from __future__ import annotations

import tarfile
import time
from collections.abc import Iterator
from contextlib import contextmanager
from pathlib import Path

import torch
import torch.profiler as tprof

from d9d_b200.core.dist_context import REGULAR_DOMAIN, DistributedContext


class Profiler:
    """``torch.profiler`` with a repeating ``wait -> warmup -> active`` schedule of length ``period_steps``.

    Every finished cycle is exported as a chrome trace, tar-gzipped to
    ``save_dir/step_{n}/rank-{r}-coord-{c0-c1-...}-trace.tar.gz`` (``trace.tar.gz`` when not distributed) and the raw
    JSON is removed.
    """

    def __init__(self, save_dir: Path, period_steps: int, warmup_steps: int, active_steps: int, dist_context: DistributedContext):
        if period_steps < warmup_steps + active_steps:
            raise ValueError("period_steps must be >= warmup_steps + active_steps")
        self._save_dir = Path(save_dir)
        self._period, self._warmup, self._active = period_steps, warmup_steps, active_steps
        self._ctx = dist_context

    def _trace_name(self) -> str:
        if not self._ctx.mesh_params.is_distributed:
            return "trace.json"
        mesh = self._ctx.mesh_for(REGULAR_DOMAIN)
        coord = mesh.get_coordinate()
        if coord is None:
            raise RuntimeError("Invalid mesh")
        return f"rank-{mesh.get_rank()}-coord-{'-'.join(map(str, coord))}-trace.json"

    def _on_trace_ready(self, prof: tprof.profile) -> None:
        out_dir = self._save_dir / f"step_{prof.step_num}"
        out_dir.mkdir(parents=True, exist_ok=True)
        raw = out_dir / self._trace_name()
        t0 = time.monotonic()
        prof.export_chrome_trace(str(raw))
        with tarfile.open(raw.with_suffix(".tar.gz"), "w:gz") as tar:
            tar.add(raw, arcname=raw.name)
        raw.unlink()
        self._ctx.logger.info(f"Finished dumping profiler traces in {time.monotonic() - t0:.2f} seconds")

    @contextmanager
    def open(self, start_step: int) -> Iterator[tprof.profile]:
        activities = [tprof.ProfilerActivity.CPU]
        if torch.cuda.is_available():
            activities.append(tprof.ProfilerActivity.CUDA)
        schedule = tprof.schedule(wait=self._period - self._warmup - self._active, warmup=self._warmup, active=self._active)
        with tprof.profile(activities=activities, schedule=schedule, on_trace_ready=self._on_trace_ready,
                           record_shapes=True, with_stack=True) as prof:
            prof.step_num = start_step
            yield prof
