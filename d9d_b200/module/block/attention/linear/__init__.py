"""Linear-attention token mixers."""

from .gated_deltanet import (
    AnyDecayGateParameters,
    CausalShortDepthwiseConv1d,
    GatedDeltaNet,
    LogSigmoidDecayGate,
    LogSigmoidDecayGateParameters,
    MambaDecayGate,
    MambaDecayGateParameters,
)

__all__ = [
    "AnyDecayGateParameters",
    "CausalShortDepthwiseConv1d",
    "GatedDeltaNet",
    "LogSigmoidDecayGate",
    "LogSigmoidDecayGateParameters",
    "MambaDecayGate",
    "MambaDecayGateParameters",
]
