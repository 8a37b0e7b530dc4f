"""Distributed context: one world, five DeviceMesh views (regular / dense / expert / batch / flat).

Parity: reference ``d9d/core/dist_context`` (``params.py:9-113``, ``configured.py:34-222``,
``device_mesh_domains.py:40-184``).  Unlike the reference the context is device-agnostic: with CUDA it uses
NCCL over NVLink/NVSwitch, without CUDA it falls back to gloo on CPU so that all plumbing is testable.
"""

from .configured import DistributedContext
from .device_mesh_domains import BATCH_DOMAIN, DENSE_DOMAIN, EXPERT_DOMAIN, FLAT_DOMAIN, REGULAR_DOMAIN
from .log import build_dist_logger
from .params import DeviceMeshParameters

__all__ = [
    "BATCH_DOMAIN",
    "DENSE_DOMAIN",
    "EXPERT_DOMAIN",
    "FLAT_DOMAIN",
    "REGULAR_DOMAIN",
    "DeviceMeshParameters",
    "DistributedContext",
    "build_dist_logger",
]
