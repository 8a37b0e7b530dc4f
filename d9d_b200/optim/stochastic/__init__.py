from .adamw import StochasticAdamW

__all__ = ["StochasticAdamW"]
