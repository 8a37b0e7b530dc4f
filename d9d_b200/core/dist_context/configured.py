from __future__ import annotations

import datetime
import logging
import os
import socket
from collections.abc import Iterator
from contextlib import contextmanager
from typing import TYPE_CHECKING

import torch
import torch.distributed as dist
from torch.distributed import DeviceMesh

from .device_mesh_domains import ALL_DOMAIN_PROVIDERS, REGULAR_DOMAIN
from .log import build_dist_logger

if TYPE_CHECKING:
    from .params import DeviceMeshParameters


def _resolve_master_addr() -> str:
    addr = os.environ.get("MASTER_ADDR")
    if addr is None:
        return "127.0.0.1"
    try:
        return socket.gethostbyname(addr)
    except OSError:
        return addr


class DistributedContext:
    """Single source of truth about the distributed environment of this process.

    * picks the device: ``cuda:{LOCAL_RANK}`` (NCCL over NVLink/NVSwitch) when CUDA is present, CPU/gloo otherwise;
    * builds the five mesh domains (all row-major views of the same rank order);
    * exposes rank / node information, a mesh-aware logger, barriers and timeout control.
    """

    def __init__(self, params: "DeviceMeshParameters", log_level: int = logging.INFO):
        self._params = params
        self._local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self._global_rank = int(os.environ.get("RANK", "0"))
        self._master_addr = _resolve_master_addr()

        use_cuda = torch.cuda.is_available()
        if use_cuda:
            torch.cuda.set_device(self._local_rank)
            self._device = torch.device("cuda", self._local_rank)
            devices_per_node = max(torch.cuda.device_count(), 1)
        else:
            self._device = torch.device("cpu")
            devices_per_node = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        self._device_type = self._device.type

        self._meshes: dict[str, DeviceMesh] = {}
        if params.is_distributed:
            if not dist.is_initialized():
                dist.init_process_group(
                    backend="nccl" if use_cuda else "gloo",
                    device_id=self._device if use_cuda else None,
                )
            if dist.get_world_size() != params.world_size:
                raise ValueError(
                    f"mesh parameters describe {params.world_size} ranks but the job has {dist.get_world_size()}"
                )
            for domain in ALL_DOMAIN_PROVIDERS:
                self._meshes[domain.name] = domain.build_mesh(params, self._device_type)
            regular = self._meshes[REGULAR_DOMAIN]
            qualifier = "-".join(
                f"{short}:{regular.get_local_rank(dim)}"
                for short, dim in (
                    ("pp", "pp"),
                    ("dpr", "dp_replicate"),
                    ("dps", "dp_shard"),
                    ("cps", "cp_shard"),
                    ("cpr", "cp_replicate"),
                    ("tp", "tp"),
                )
            )
            self._num_nodes = max(regular.size() // devices_per_node, 1)
        else:
            qualifier = "local"
            self._num_nodes = 1
        self._node_rank = self._global_rank // devices_per_node
        self._logger = build_dist_logger(qualifier, level=log_level)

    # ---------------------------------------------------------------- info
    @property
    def logger(self) -> logging.Logger:
        return self._logger

    @property
    def mesh_params(self) -> "DeviceMeshParameters":
        return self._params

    @property
    def is_distributed(self) -> bool:
        return self._params.is_distributed

    @property
    def current_device(self) -> torch.device:
        return self._device

    @property
    def device_type(self) -> str:
        return self._device_type

    @property
    def is_main_process(self) -> bool:
        return self._global_rank == 0

    @property
    def is_local_main_process(self) -> bool:
        return self._local_rank == 0

    @property
    def global_rank(self) -> int:
        return self._global_rank

    @property
    def local_rank(self) -> int:
        return self._local_rank

    @property
    def node_rank(self) -> int:
        return self._node_rank

    @property
    def num_nodes(self) -> int:
        return self._num_nodes

    @property
    def master_addr(self) -> str:
        return self._master_addr

    def mesh_for(self, domain: str) -> DeviceMesh:
        """Mesh view for ``regular | dense | expert | batch | flat``."""
        try:
            return self._meshes[domain]
        except KeyError:
            raise ValueError(f"Domain {domain} does not exist") from None

    # ---------------------------------------------------------------- synchronisation
    def wait_world(self) -> None:
        """Barrier over the world followed by a device sync."""
        if self._params.is_distributed:
            if self._device_type == "cuda":
                dist.barrier(device_ids=[torch.cuda.current_device()])
            else:
                dist.barrier()
        if self._device_type == "cuda":
            torch.cuda.synchronize()

    def set_timeout(self, timeout_seconds: float) -> None:
        """Set the collective timeout on the world group and every mesh-dim group."""
        if not self._params.is_distributed:
            return
        self.logger.info(f"Setting global timeout to {timeout_seconds} seconds")
        self.wait_world()
        delta = datetime.timedelta(seconds=timeout_seconds)
        seen: set[int] = set()
        groups: list[dist.ProcessGroup | None] = [None]
        for mesh in self._meshes.values():
            for dim in range(mesh.ndim):
                group = mesh.get_group(dim)
                if id(group) not in seen:
                    seen.add(id(group))
                    groups.append(group)
        for group in groups:
            try:
                dist.distributed_c10d._set_pg_timeout(delta, group)  # noqa: SLF001
            except Exception as exc:  # gloo backends do not always expose a mutable timeout
                self.logger.debug(f"could not set timeout on {group}: {exc}")

    @contextmanager
    def local_main_process_first(self) -> Iterator[None]:
        if not self.is_local_main_process:
            self.wait_world()
        yield
        if self.is_local_main_process:
            self.wait_world()

    @contextmanager
    def main_process_first(self) -> Iterator[None]:
        if not self.is_main_process:
            self.wait_world()
        yield
        if self.is_main_process:
            self.wait_world()
