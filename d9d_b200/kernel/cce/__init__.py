from .main import LinearCrossEntropy, linear_cross_entropy

__all__ = ["LinearCrossEntropy", "linear_cross_entropy"]
