from .flash import FlashSdpa

__all__ = ["FlashSdpa"]
