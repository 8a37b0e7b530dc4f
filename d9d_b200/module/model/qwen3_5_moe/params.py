"""Hyper-parameters of the Qwen3.5-MoE (text) family: the hybrid Gated DeltaNet / gated attention stack of ``qwen3_5`` with a
mixture-of-experts feed-forward (softmax top-k router, renormalised) plus a sigmoid-gated shared expert in every layer."""

from pydantic import PositiveInt, model_validator

from d9d_b200.module.model._params import family_parameters
from d9d_b200.module.model.qwen3_5.params import HybridMixerFields


class HybridMoELayerFields(HybridMixerFields):
    moe_intermediate_size: PositiveInt
    shared_expert_intermediate_size: PositiveInt
    num_experts: PositiveInt
    experts_top_k: PositiveInt

    @model_validator(mode="after")
    def _check_top_k(self):  # noqa: ANN202
        if self.experts_top_k > self.num_experts:
            raise ValueError("experts_top_k cannot exceed num_experts")
        return self


_generated = family_parameters("Qwen3_5MoE", HybridMoELayerFields, __name__)

Qwen3_5MoELayerParameters = _generated["Qwen3_5MoELayerParameters"]
Qwen3_5MoEParameters = _generated["Qwen3_5MoEParameters"]
Qwen3_5MoEForCausalLMParameters = _generated["Qwen3_5MoEForCausalLMParameters"]
Qwen3_5MoEForClassificationParameters = _generated["Qwen3_5MoEForClassificationParameters"]
Qwen3_5MoEForEmbeddingParameters = _generated["Qwen3_5MoEForEmbeddingParameters"]

__all__ = list(_generated)
