"""Streaming HF-style safetensors IO driven by mapper dependency groups (reference ``d9d/model_state/io``)."""

from .module_reader import load_model_state
from .module_writer import save_model_state, save_model_state_pipeline_parallel
from .reader import read_model_state
from .writer import write_model_state_distributed, write_model_state_local, write_model_state_pipeline_parallel

__all__ = [
    "load_model_state",
    "read_model_state",
    "save_model_state",
    "save_model_state_pipeline_parallel",
    "write_model_state_distributed",
    "write_model_state_local",
    "write_model_state_pipeline_parallel",
]
