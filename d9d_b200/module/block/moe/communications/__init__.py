"""Token dispatch / combine strategies for MoE layers."""

from .base import ExpertCommunicationHandler
from .naive import NoCommunicationHandler

__all__ = ["ExpertCommunicationHandler", "NoCommunicationHandler"]
