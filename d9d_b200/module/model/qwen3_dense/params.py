"""Hyper-parameters of the Qwen3Dense family (classes generated from the shared field sets in ``module/model/_params.py``)."""

from d9d_b200.module.model._params import DenseLayerFields, family_parameters

_generated = family_parameters("Qwen3Dense", DenseLayerFields, __name__)

Qwen3DenseLayerParameters = _generated["Qwen3DenseLayerParameters"]
Qwen3DenseParameters = _generated["Qwen3DenseParameters"]
Qwen3DenseForCausalLMParameters = _generated["Qwen3DenseForCausalLMParameters"]
Qwen3DenseForClassificationParameters = _generated["Qwen3DenseForClassificationParameters"]
Qwen3DenseForEmbeddingParameters = _generated["Qwen3DenseForEmbeddingParameters"]

__all__ = list(_generated)
