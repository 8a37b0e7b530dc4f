"""DeepSeek-V3 = the DeepSeek-V2 architecture (``module/model/deepseek_v2``) with ``layer.router`` set to sigmoid scores, a
selection-bias buffer, group-limited top-k and weight scaling.  This package offers the same classes and mappers under V3
names; ``deepseek_v3_router(...)`` builds the router options from the HuggingFace config fields."""
from d9d_b200.module.block.moe.router import RouterParameters
from d9d_b200.module.model import deepseek_v2 as _v2

DeepseekV3ExpertsFormat = _v2.DeepseekV2ExpertsFormat
DeepseekV3ForCausalLM = _v2.DeepseekV2ForCausalLM
DeepseekV3ForCausalLMParameters = _v2.DeepseekV2ForCausalLMParameters
DeepseekV3ForClassification = _v2.DeepseekV2ForClassification
DeepseekV3ForClassificationParameters = _v2.DeepseekV2ForClassificationParameters
DeepseekV3ForEmbedding = _v2.DeepseekV2ForEmbedding
DeepseekV3ForEmbeddingParameters = _v2.DeepseekV2ForEmbeddingParameters
DeepseekV3Layer = _v2.DeepseekV2Layer
DeepseekV3LayerParameters = _v2.DeepseekV2LayerParameters
DeepseekV3Model = _v2.DeepseekV2Model
DeepseekV3Parameters = _v2.DeepseekV2Parameters
mapper_from_huggingface_deepseek_v3 = _v2.mapper_from_huggingface_deepseek_v2
mapper_from_huggingface_deepseek_v3_for_causal_lm = _v2.mapper_from_huggingface_deepseek_v2_for_causal_lm
mapper_from_huggingface_deepseek_v3_for_classification = _v2.mapper_from_huggingface_deepseek_v2_for_classification
mapper_from_huggingface_deepseek_v3_for_embedding = _v2.mapper_from_huggingface_deepseek_v2_for_embedding
mapper_to_huggingface_deepseek_v3 = _v2.mapper_to_huggingface_deepseek_v2
mapper_to_huggingface_deepseek_v3_for_causal_lm = _v2.mapper_to_huggingface_deepseek_v2_for_causal_lm
mapper_to_huggingface_deepseek_v3_for_classification = _v2.mapper_to_huggingface_deepseek_v2_for_classification
mapper_to_huggingface_deepseek_v3_for_embedding = _v2.mapper_to_huggingface_deepseek_v2_for_embedding


def deepseek_v3_router(n_group: int, topk_group: int, routed_scaling_factor: float) -> RouterParameters:
    """Router options of DeepSeek-V3 from the HuggingFace config fields of the same names."""
    return RouterParameters(score_function="sigmoid", enable_expert_bias=True, num_expert_groups=n_group,
                            topk_expert_groups=topk_group, routed_scaling_factor=routed_scaling_factor)


__all__ = [
    "DeepseekV3ExpertsFormat",
    "DeepseekV3ForCausalLM",
    "DeepseekV3ForCausalLMParameters",
    "DeepseekV3ForClassification",
    "DeepseekV3ForClassificationParameters",
    "DeepseekV3ForEmbedding",
    "DeepseekV3ForEmbeddingParameters",
    "DeepseekV3Layer",
    "DeepseekV3LayerParameters",
    "DeepseekV3Model",
    "DeepseekV3Parameters",
    "deepseek_v3_router",
    "mapper_from_huggingface_deepseek_v3",
    "mapper_from_huggingface_deepseek_v3_for_causal_lm",
    "mapper_from_huggingface_deepseek_v3_for_classification",
    "mapper_from_huggingface_deepseek_v3_for_embedding",
    "mapper_to_huggingface_deepseek_v3",
    "mapper_to_huggingface_deepseek_v3_for_causal_lm",
    "mapper_to_huggingface_deepseek_v3_for_classification",
    "mapper_to_huggingface_deepseek_v3_for_embedding",
]
