from .rms_norm import RMSNorm

__all__ = ["RMSNorm"]
