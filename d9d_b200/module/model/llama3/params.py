"""Hyper-parameters of the Llama3 family (classes generated from the shared field sets in ``module/model/_params.py``)."""

from d9d_b200.module.model._params import DenseLayerFields, family_parameters

_generated = family_parameters("Llama3", DenseLayerFields, __name__)

Llama3LayerParameters = _generated["Llama3LayerParameters"]
Llama3Parameters = _generated["Llama3Parameters"]
Llama3ForCausalLMParameters = _generated["Llama3ForCausalLMParameters"]
Llama3ForClassificationParameters = _generated["Llama3ForClassificationParameters"]
Llama3ForEmbeddingParameters = _generated["Llama3ForEmbeddingParameters"]

__all__ = list(_generated)
