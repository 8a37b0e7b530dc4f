"""Rounding primitives shared by the stochastic kernels (reference ``d9d/kernel/stochastic/ops/round.py:5-24``).

In the reference ``fp32_to_bf16_kernel`` is a Triton device function inlined into the AdamW / copy kernels.  Here the
device-side twin is ``sr_round_bf16`` in ``ops/csrc/elementwise.cu`` (Philox keyed by ``(seed, element offset)``); this
module exposes the same primitive as a tensor-level function so that custom optimizers can reuse it.
"""

from .round import fp32_to_bf16_kernel

__all__ = ["fp32_to_bf16_kernel"]
