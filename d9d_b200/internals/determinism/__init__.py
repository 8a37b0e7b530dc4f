"""Seeding (reference ``internals/determinism/seed.py:11-61``)."""

from .seed import set_seeds

__all__ = ["set_seeds"]
