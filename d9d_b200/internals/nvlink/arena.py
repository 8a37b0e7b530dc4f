from __future__ import annotations

import torch
import torch.distributed as dist


def nvlink_available() -> bool:
    """True when CUDA peer memory through ``torch.distributed._symmetric_memory`` can be used."""
    if not torch.cuda.is_available() or not dist.is_available() or not dist.is_initialized():
        return False
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
    except Exception:
        return False
    return True


class SymmetricArena:
    """A flat buffer with the same size on every rank of ``group``, mapped into every peer's address space.

    * ``buffer``         this rank's memory (an ordinary CUDA tensor for local kernels),
    * ``peer_ptrs_dev``  device address of a ``[world]`` table with every replica's base pointer (P2P loads/stores),
    * ``multicast_ptr``  NVSwitch multicast (NVLS) mapping of all replicas or 0 — ``multimem.ld_reduce`` through it
                         sums the replicas inside the switch, ``multimem.st`` writes all replicas with one store,
    * ``barrier()``      device-side barrier over the group on the current stream (signal pads, no host sync).

    Allocation and rendezvous go through ``torch.distributed._symmetric_memory`` (CUDA VMM + handle exchange); the
    kernels that use the pointers are ours (``ops/csrc/nvlink_optim.cu``).
    """

    def __init__(self, numel: int, dtype: torch.dtype, device: torch.device, group: dist.ProcessGroup):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group
        self.buffer = symm_mem.empty(numel, dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.buffer, group.group_name)
        self.world_size: int = self.handle.world_size
        self.rank: int = self.handle.rank
        self.peer_ptrs_dev: int = int(self.handle.buffer_ptrs_dev)
        self.multicast_ptr: int = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self._barrier_channel = 0

    def barrier(self) -> None:
        self.handle.barrier(channel=self._barrier_channel)

    def peer_view(self, rank: int, shape: tuple[int, ...] | None = None) -> torch.Tensor:
        """Tensor aliasing replica ``rank``'s buffer (debugging / tests)."""
        shape = tuple(self.buffer.shape) if shape is None else shape
        return self.handle.get_buffer(rank, shape, self.buffer.dtype)
