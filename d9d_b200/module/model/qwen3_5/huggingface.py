"""HuggingFace ⇄ native state mappers of the Qwen3.5 text family (rules in ``module/model/_huggingface.py``).

HF names the token mixer of a layer ``linear_attn`` or ``self_attn`` depending on its type; natively it is always
``self_attn``.  The full-attention query projection is fused with the output gate head by head in HF."""

from __future__ import annotations

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.module.model import _huggingface as hf

from .params import (
    Qwen3_5ForCausalLMParameters,
    Qwen3_5ForClassificationParameters,
    Qwen3_5ForEmbeddingParameters,
    Qwen3_5LayerParameters,
    Qwen3_5Parameters,
)

_LINEAR_RENAMES = (("in_proj_qkv.weight", "qkv_proj.weight"), ("in_proj_z.weight", "g_proj.weight"), ("in_proj_b.weight", "b_proj.weight"),
                   ("in_proj_a.weight", "decay_gate.proj.weight"), ("A_log", "decay_gate.A_log"), ("dt_bias", "decay_gate.dt_bias"),
                   ("norm.weight", "out_norm.weight"), ("out_proj.weight", "o_proj.weight"))


def _token_mixer(layer: Qwen3_5LayerParameters, index: int) -> tuple[hf.Rule, ...]:
    if layer.uses_full_attention(index):
        return (hf.HeadInterleaved("self_attn.q_proj.weight", ("self_attn.q_proj.weight", "self_attn.gate_proj.weight"), layer.head_dim),
                *(hf.Same(f"self_attn.{name}.weight") for name in ("k_proj", "v_proj", "o_proj", "q_norm", "k_norm")))
    return (*(hf.Renamed(f"linear_attn.{a}", f"self_attn.{b}") for a, b in _LINEAR_RENAMES),
            hf.Unsqueezed("linear_attn.conv1d.weight", "self_attn.qkv_conv1d.weight", dim=1))


def _backbone(params: Qwen3_5Parameters) -> tuple[hf.Rule, ...]:
    def layer_rules(index: int) -> tuple[hf.Rule, ...]:
        return (*_token_mixer(params.layer, index), *hf.norm_rules(), *hf.dense_mlp_rules())

    return hf.backbone_rules(layer_rules, params.num_hidden_layers, hf.single_vocab_name(params.split_vocab_order))


def mapper_from_huggingface_qwen3_5(params: Qwen3_5Parameters) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_5(params: Qwen3_5Parameters) -> ModelStateMapper:
    return hf.compile_rules(_backbone(params), hf.Direction.TO_HF)


def _causal(params: Qwen3_5ForCausalLMParameters) -> tuple[hf.Rule, ...]:
    return hf.causal_lm_rules(_backbone(params.model), hf.single_vocab_name(params.model.split_vocab_order))


def mapper_from_huggingface_qwen3_5_for_causal_lm(params: Qwen3_5ForCausalLMParameters) -> ModelStateMapper:
    return hf.compile_rules(_causal(params), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_5_for_causal_lm(params: Qwen3_5ForCausalLMParameters) -> ModelStateMapper:
    return hf.compile_rules(_causal(params), hf.Direction.TO_HF)


def mapper_from_huggingface_qwen3_5_for_classification(params: Qwen3_5ForClassificationParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model)), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_5_for_classification(params: Qwen3_5ForClassificationParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.classification_rules(_backbone(params.model)), hf.Direction.TO_HF)


def mapper_from_huggingface_qwen3_5_for_embedding(params: Qwen3_5ForEmbeddingParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model)), hf.Direction.FROM_HF)


def mapper_to_huggingface_qwen3_5_for_embedding(params: Qwen3_5ForEmbeddingParameters) -> ModelStateMapper:
    return hf.compile_rules(hf.embedding_rules(_backbone(params.model)), hf.Direction.TO_HF)
