from .classification import ClassificationHead
from .embedding import EmbeddingHead
from .language_modelling import LM_IGNORE_INDEX, SplitLanguageModellingHead

__all__ = ["LM_IGNORE_INDEX", "ClassificationHead", "EmbeddingHead", "SplitLanguageModellingHead"]
