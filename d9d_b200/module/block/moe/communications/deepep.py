"""Import-path compatibility for code written against the reference's DeepEP handler.

The reference moves tokens between expert-parallel ranks with the external ``deep_ep`` library
(``d9d/module/block/moe/communications/deepep.py:14-222``: NVLink / NVSHMEM buffers, a host wait for the receive
counts).  This framework has no such dependency: the same role is filled by our own NVLink peer-memory kernels
(:mod:`.nvlink` - count exchange, destination rows computed on the device, push into GEMM-ready buffers, pull-sum
combine, no host synchronisation) with the NCCL all-to-all handler (:mod:`.expert_parallel`) as the fallback.
The names below let ``from ...communications.deepep import DeepEpCommunicationHandler`` keep working.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from .nvlink import AutoExpertParallelCommunicationHandler, _Workspace

DeepEpCommunicationHandler = AutoExpertParallelCommunicationHandler


def get_hidden_state_bytes(x: torch.Tensor) -> int:
    """Bytes one token occupies in the dispatch buffers (at least bf16-sized, like the reference's ``:14-22``)."""
    return x.shape[-1] * max(x.element_size(), 2)


def init_deepep_buffer(group: dist.ProcessGroup, hidden_bytes: int) -> _Workspace:
    """Return the symmetric-memory workspace shared by every MoE layer of ``group`` (reference ``:25-53``).

    The workspace is sized lazily by the first dispatch (it needs the token count), so ``hidden_bytes`` only has to be
    accepted here; calling this early is harmless and allocation-free.
    """
    del hidden_bytes
    return _Workspace.for_group(group)


__all__ = ["DeepEpCommunicationHandler", "get_hidden_state_bytes", "init_deepep_buffer"]
