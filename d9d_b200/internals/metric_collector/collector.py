from __future__ import annotations

from typing import TYPE_CHECKING, Any

import torch
import torch.utils._pytree as pytree
from torch.profiler import record_function

from d9d_b200.core.dist_context import DistributedContext

if TYPE_CHECKING:
    from d9d_b200.metric import Metric


class AsyncMetricCollector:
    """Runs ``metric.sync()`` + ``metric.compute()`` off the critical path.

    On CUDA the work is queued on a dedicated stream that first waits for the training stream; results are joined,
    copied to the host and converted to python scalars only in :meth:`collect_results`.  On CPU it degrades to
    synchronous execution (same API).
    """

    def __init__(self, metric: "Metric"):
        self._metric = metric
        self._stream: "torch.cuda.Stream | None" = None
        self._bound = False
        self._pending: Any = None

    def bind(self, device: torch.device | str | None = None) -> None:
        device = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self._metric.to(device)
        self._stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        self._bound = True

    def unbind(self) -> None:
        self._stream = None
        self._bound = False

    def schedule_collection(self, dist_context: DistributedContext) -> None:
        if not self._bound:
            raise RuntimeError("AsyncMetricSynchronizer is not bound. Call .bind() first.")
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream), record_function("Async Metric Sync & Compute"):
                if dist_context.mesh_params.is_distributed:
                    self._metric.sync(dist_context)
                self._pending = self._metric.compute()
        else:
            if dist_context.mesh_params.is_distributed:
                self._metric.sync(dist_context)
            self._pending = self._metric.compute()

    def collect_results(self) -> Any:
        if not self._bound:
            raise RuntimeError("AsyncMetricSynchronizer is not bound. Call .bind() first.")
        if self._pending is None:
            raise RuntimeError("sync_and_compute() was not called.")
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        results, self._pending = self._pending, None
        results = pytree.tree_map(lambda x: x.cpu().item() if isinstance(x, torch.Tensor) else x, results)
        self._metric.reset()
        return results
