"""Small shared helpers re-exported in one place."""

from d9d_b200.internals.determinism import set_seeds
from d9d_b200.kernel._native import native_launch_count, native_ops

__all__ = ["native_launch_count", "native_ops", "set_seeds"]
