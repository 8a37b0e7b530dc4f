"""NVLink / NVSwitch peer-memory substrate: symmetric arenas every replica can load from / store to directly."""

from .arena import SymmetricArena, nvlink_available

__all__ = ["SymmetricArena", "nvlink_available"]
