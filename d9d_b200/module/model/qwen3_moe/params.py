"""Hyper-parameters of the Qwen3MoE family (classes generated from the shared field sets in ``module/model/_params.py``)."""

from d9d_b200.module.model._params import MoELayerFields, family_parameters

_generated = family_parameters("Qwen3MoE", MoELayerFields, __name__)

Qwen3MoELayerParameters = _generated["Qwen3MoELayerParameters"]
Qwen3MoEParameters = _generated["Qwen3MoEParameters"]
Qwen3MoEForCausalLMParameters = _generated["Qwen3MoEForCausalLMParameters"]
Qwen3MoEForClassificationParameters = _generated["Qwen3MoEForClassificationParameters"]
Qwen3MoEForEmbeddingParameters = _generated["Qwen3MoEForEmbeddingParameters"]

__all__ = list(_generated)
