"""Full fine-tuning of selected sub-modules expressed as a PEFT method."""

from .config import FullTuneConfig
from .method import FullTune

__all__ = ["FullTune", "FullTuneConfig"]
