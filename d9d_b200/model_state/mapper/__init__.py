"""Core of the state-mapping system: a mapper *declares* its dependency groups and *applies* them on tensors."""

from .abc import ModelStateMapper, StateGroup

__all__ = ["ModelStateMapper", "StateGroup"]
