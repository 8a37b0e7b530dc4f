"""FSDP parameter all-gather and gradient reduce-scatter over NVLink peer memory instead of NCCL (opt-in).

FSDP2 lets a module replace its two collectives (``FSDPModule.set_custom_all_gather / set_custom_reduce_scatter``).  The
implementations here keep FSDP's hooks, streams and DTensor placements untouched and only change how the bytes move:

* every rank copies what it contributes (its parameter shard / its unsharded gradients) into a *symmetric* staging buffer
  that is mapped into every peer's address space (``internals.nvlink.SymmetricArena``),
* after a device-side barrier each rank **pulls** what it needs straight out of its peers' staging buffers with plain
  loads over NVLink / NVSwitch (all-gather: the other ranks' shards; reduce-scatter: its own slice of everybody's gradients,
  summed in fp32), and a second barrier releases the staging buffers for the next call.

No NCCL kernel runs for the two collectives and nothing is synchronised with the host.  The staging buffer of a process group
is shared by all FSDP units of that group (their collectives are serialised on FSDP's communication streams anyway) and
grows on demand (a collective allocation, only during the first step).

Status: selected with ``D9D_FSDP_COMM=peer`` (see ``parallelize_fsdp``); the protocol is exercised on CPU / gloo through an
emulated arena (``tests/test_fsdp_peer_comm_gloo.py``); it has not been timed on GPUs yet, which is why NCCL stays the default.
This is the staging-copy variant - fusing the pull into the first GEMM that consumes the weights (arrival flags per shard, as
the tensor-parallel ``COMM_WAIT_A`` GEMM does for activations) is the next step.
"""

from __future__ import annotations

from collections.abc import Callable, Sequence
from typing import Any

import torch
import torch.distributed as dist
from torch.distributed.fsdp._fully_shard._fsdp_api import AllGather, ReduceScatter


class _PeerStage:
    """Growable symmetric byte buffer of one process group."""

    def __init__(self, group: dist.ProcessGroup, device: torch.device, arena_factory: Callable[[int, torch.device, dist.ProcessGroup], Any]):
        self._group, self._device, self._factory = group, device, arena_factory
        self._arena: Any = None
        self._capacity = 0

    def ensure(self, nbytes: int) -> None:
        if nbytes <= self._capacity:
            return
        capacity = 1 << max(20, (nbytes - 1).bit_length())  # power of two, at least 1 MiB: few re-allocations
        if self._arena is not None:
            self._arena.barrier()  # nobody still reads the old buffer
        self._arena = self._factory(capacity, self._device, self._group)
        self._capacity = capacity

    def local(self, numel: int, dtype: torch.dtype) -> torch.Tensor:
        return self._arena.buffer[: numel * dtype.itemsize].view(dtype)

    def peers(self, numel: int, dtype: torch.dtype) -> list[torch.Tensor]:
        """One tensor per rank aliasing (or, for the emulated arena, holding a snapshot of) that rank's staging buffer."""
        world = self._group.size()
        if hasattr(self._arena, "snapshot"):
            return [t[: numel * dtype.itemsize].view(dtype) for t in self._arena.snapshot()]
        return [self._arena.peer_view(r, (self._capacity,))[: numel * dtype.itemsize].view(dtype) for r in range(world)]

    def barrier(self) -> None:
        self._arena.barrier()


def _default_arena(nbytes: int, device: torch.device, group: dist.ProcessGroup):
    from d9d_b200.internals.nvlink import SymmetricArena

    return SymmetricArena(nbytes, torch.uint8, device, group)


class _PeerComm:
    def __init__(self, stage: _PeerStage):
        self._stage = stage

    def allocate(self, size: Sequence[int | torch.SymInt], *, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
        # outputs become long-lived tensors (unsharded parameters are copied out of them, sharded gradients are views of
        # them): ordinary allocations; only the *staged* copy of a collective's input lives in peer-visible memory
        return torch.empty(*size, dtype=dtype, device=device)


class PeerMemoryAllGather(_PeerComm, AllGather):
    def __call__(self, output_tensor: torch.Tensor, input_tensor: torch.Tensor, group: dist.ProcessGroup, async_op: bool = False):
        world, rank, n = group.size(), group.rank(), input_tensor.numel()
        if output_tensor.numel() != n * world:
            raise ValueError("all-gather output must hold one input per rank")
        stage = self._stage
        stage.ensure(n * input_tensor.element_size())
        stage.local(n, input_tensor.dtype).copy_(input_tensor.reshape(-1))
        stage.barrier()  # every rank's shard is in its staging buffer
        out = output_tensor.reshape(-1)
        peers = stage.peers(n, input_tensor.dtype)
        for p in range(world):
            dst = out[p * n : (p + 1) * n]
            if p == rank:
                if dst.data_ptr() != input_tensor.data_ptr():  # FSDP hands the input in as this very slice of the output
                    dst.copy_(input_tensor.reshape(-1))
            else:
                dst.copy_(peers[p])  # peer-to-peer load over NVLink
        stage.barrier()  # all pulls are done: the staging buffers may be overwritten by the next collective
        return None


class PeerMemoryReduceScatter(_PeerComm, ReduceScatter):
    def __call__(self, output_tensor: torch.Tensor, input_tensor: torch.Tensor, group: dist.ProcessGroup, op: Any, async_op: bool = False):
        world, rank = group.size(), group.rank()
        total, n = input_tensor.numel(), output_tensor.numel()
        if total != n * world:
            raise ValueError("reduce-scatter input must hold one output-sized chunk per rank")
        if op == dist.ReduceOp.SUM:
            divide = 1.0
        elif op == dist.ReduceOp.AVG:
            divide = float(world)
        else:
            raise NotImplementedError(f"peer-memory reduce-scatter implements SUM / AVG, got {op}")
        stage = self._stage
        stage.ensure(total * input_tensor.element_size())
        flat = input_tensor.reshape(-1)
        stage.local(total, input_tensor.dtype).copy_(flat)
        stage.barrier()  # every rank's gradients are staged
        peers = stage.peers(total, input_tensor.dtype)
        acc = flat[rank * n : (rank + 1) * n].float()
        for p in range(world):
            if p != rank:
                acc += peers[p][rank * n : (rank + 1) * n].float()  # my slice of rank p's gradients, pulled over NVLink
        if divide != 1.0:
            acc /= divide
        output_tensor.reshape(-1).copy_(acc)
        stage.barrier()
        return None


_COMMS: dict[tuple[str, str, int], tuple[PeerMemoryAllGather, PeerMemoryReduceScatter]] = {}


def peer_memory_comms(group: dist.ProcessGroup, device: torch.device,
                      arena_factory: Callable[[int, torch.device, dist.ProcessGroup], Any] | None = None,
                      ) -> tuple[PeerMemoryAllGather, PeerMemoryReduceScatter]:
    """The (all-gather, reduce-scatter) pair of ``group`` - one per (group, device, arena factory), shared by every FSDP unit
    sharded over the group.  The two have their own staging buffers because FSDP runs them on different streams."""
    factory = arena_factory or _default_arena
    key = (group.group_name, str(device), id(factory))
    if key not in _COMMS:
        _COMMS[key] = (PeerMemoryAllGather(_PeerStage(group, device, factory)), PeerMemoryReduceScatter(_PeerStage(group, device, factory)))
    return _COMMS[key]
