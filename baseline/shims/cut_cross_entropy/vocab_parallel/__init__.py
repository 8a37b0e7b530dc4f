from .utils import VocabParallelOptions, vp_reduce_correct_logit, vp_reduce_lse  # noqa: F401
