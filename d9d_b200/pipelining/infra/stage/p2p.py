"""Point-to-point transport of a stage boundary (activations forward, gradients backward).

One :class:`BoundaryChannel` per direction and stage.  Receive buffers are allocated when the RECV action is issued
(caching-allocator alloc, no device sync) and handed to the consumer; send/recv op lists are sorted by tensor name
on both sides so both ends agree on the order inside a batched NCCL group.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


class BoundaryChannel:
    def __init__(self, name: str, stage_index: int, peer_rank_in_group: int | None, group: dist.ProcessGroup | None,
                 recv_meta: dict[str, torch.Tensor], device: torch.device, requires_grad: bool):
        self._name = name
        self._stage = stage_index
        self._peer = peer_rank_in_group
        self._group = group
        self._meta = dict(sorted(recv_meta.items()))
        self._device = device
        self._requires_grad = requires_grad
        self._received: dict[int, dict[str, torch.Tensor]] = {}

    # -------------------------------------------------------------- receive side
    def make_recv_ops(self, microbatch: int) -> list[dist.P2POp]:
        if self._peer is None or self._group is None:
            return []
        bufs = {k: torch.empty(m.shape, dtype=m.dtype, device=self._device) for k, m in self._meta.items()}
        self._received[microbatch] = bufs
        return [dist.P2POp(dist.irecv, buf, group=self._group, group_peer=self._peer) for buf in bufs.values()]

    def set_local(self, tensors: dict[str, torch.Tensor], microbatch: int) -> None:
        """Same-rank hand-off from the neighbouring stage (no copy)."""
        self._received[microbatch] = {k: v.detach() for k, v in tensors.items()}

    def take(self, microbatch: int) -> dict[str, torch.Tensor]:
        tensors = self._received.pop(microbatch)
        if self._requires_grad:
            for t in tensors.values():
                if t.is_floating_point():
                    t.requires_grad_(True)
        return tensors

    # -------------------------------------------------------------- send side
    def make_send_ops(self, tensors: dict[str, torch.Tensor]) -> list[dist.P2POp]:
        if self._peer is None or self._group is None:
            return []
        ops = []
        for _, t in sorted(tensors.items()):
            ops.append(dist.P2POp(dist.isend, t.detach().contiguous(), group=self._group, group_peer=self._peer))
        return ops

    def pending(self) -> int:
        return len(self._received)

    def reset(self) -> None:
        self._received.clear()
