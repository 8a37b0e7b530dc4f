"""Hyper-parameters of the Mixtral family (classes generated from the shared field sets in ``module/model/_params.py``)."""

from d9d_b200.module.model._params import MoELayerFields, family_parameters

_generated = family_parameters("Mixtral", MoELayerFields, __name__)

MixtralLayerParameters = _generated["MixtralLayerParameters"]
MixtralParameters = _generated["MixtralParameters"]
MixtralForCausalLMParameters = _generated["MixtralForCausalLMParameters"]
MixtralForClassificationParameters = _generated["MixtralForClassificationParameters"]
MixtralForEmbeddingParameters = _generated["MixtralForEmbeddingParameters"]

__all__ = list(_generated)
