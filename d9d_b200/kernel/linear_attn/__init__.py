"""Linear-attention primitives used by ``GatedDeltaNet``: the gated delta rule, the causal short convolution and the
Mamba-style decay gate (the reference takes these from the ``fla`` Triton package —
``d9d/module/block/attention/linear/gated_deltanet.py:6-8``)."""

from .conv import causal_conv1d_silu
from .delta_rule import chunk_gated_delta_rule, recurrent_gated_delta_rule
from .gate import mamba_decay_gate

__all__ = ["causal_conv1d_silu", "chunk_gated_delta_rule", "mamba_decay_gate", "recurrent_gated_delta_rule"]
