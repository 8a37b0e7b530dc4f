"""Periodic torch-profiler traces named after the rank's mesh coordinates (reference ``internals/profiling``)."""

from .profile import Profiler

__all__ = ["Profiler"]
