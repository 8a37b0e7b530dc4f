from .stage import PipelineStage

__all__ = ["PipelineStage"]
