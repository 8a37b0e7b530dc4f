"""Small dtype helpers shared by kernels (reference ``d9d/kernel/general``)."""

from .get_int_dtype import get_int_dtype

__all__ = ["get_int_dtype"]
