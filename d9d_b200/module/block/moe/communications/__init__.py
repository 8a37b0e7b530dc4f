"""Token dispatch / combine strategies for MoE layers."""

from .base import ExpertCommunicationHandler
from .expert_parallel import ExpertParallelCommunicationHandler
from .naive import NoCommunicationHandler

__all__ = ["ExpertCommunicationHandler", "ExpertParallelCommunicationHandler", "NoCommunicationHandler"]
