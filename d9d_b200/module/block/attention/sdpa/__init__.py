from .flash import ContextParallelMode, FlashSdpa

__all__ = ["ContextParallelMode", "FlashSdpa"]
