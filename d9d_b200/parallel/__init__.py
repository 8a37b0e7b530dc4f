"""Parallelisation entry points (alias of :mod:`d9d_b200.module.parallelism` and the pipelining / NVLink layers)."""

from d9d_b200.module.parallelism.api import *  # noqa: F401,F403
from d9d_b200.module.parallelism import api, model, style  # noqa: F401
