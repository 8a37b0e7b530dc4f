"""Process-global switch telling custom autograd Functions which gradient edges to compute.

Needed for split backward (dI / dW) in zero-bubble pipeline schedules: ``ctx.needs_input_grad`` is True for every
edge that requires grad even when ``torch.autograd.backward(inputs=...)`` restricts the pass.
Parity: reference ``d9d/core/autograd/grad_context.py:5-85``.
"""

from .grad_context import GLOBAL_GRAD_CONTEXT, GlobalGradContext, GradDirection

__all__ = ["GLOBAL_GRAD_CONTEXT", "GlobalGradContext", "GradDirection"]
