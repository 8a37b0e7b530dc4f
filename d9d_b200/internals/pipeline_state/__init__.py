"""Per-step scratch state shared between task callbacks, viewable globally or per pipeline microbatch.

Parity: reference ``d9d/internals/pipeline_state`` (``handler.py:61-110``, ``storage.py:38-227``).
"""

from .api import PipelineState
from .handler import PipelineStateHandler

__all__ = ["PipelineState", "PipelineStateHandler"]
