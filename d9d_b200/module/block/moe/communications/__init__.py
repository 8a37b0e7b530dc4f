"""Token dispatch / combine strategies for MoE layers."""

from .base import ExpertCommunicationHandler
from .expert_parallel import ExpertParallelCommunicationHandler
from .naive import NoCommunicationHandler
from .nvlink import AutoExpertParallelCommunicationHandler, NvlinkExpertParallelCommunicationHandler

__all__ = ["ExpertCommunicationHandler", "ExpertParallelCommunicationHandler", "AutoExpertParallelCommunicationHandler", "NoCommunicationHandler", "NvlinkExpertParallelCommunicationHandler"]
