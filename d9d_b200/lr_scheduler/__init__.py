"""Learning-rate scheduling utilities (reference ``d9d/lr_scheduler``)."""

from .visualizer import lr_history, visualize_lr_scheduler

__all__ = ["lr_history", "visualize_lr_scheduler"]
