"""HF <-> native mappers: round trip over every family / head / expert layout, and numerical parity with the
transformers implementation where it is importable (CPU, fp32, tiny configs)."""

import json

import pytest
import torch

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
from d9d_b200.pipelining.api import PipelineStageInfo


def _run(mapper: ModelStateMapper, state: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    out = {}
    used = set()
    for group in mapper.state_dependency_groups():
        assert group.inputs <= state.keys(), sorted(group.inputs - state.keys())[:3]
        out.update(mapper.apply({k: state[k] for k in group.inputs}))
        used |= group.inputs
    assert used == set(state), sorted(set(state) - used)[:5]
    return out


VOCAB = dict(split_vocab_size={"text": 96}, split_vocab_order=["text"])


def _moe_model(family):
    if family == "qwen3_moe":
        from d9d_b200.module.model.qwen3_moe import (Qwen3MoEForCausalLM as M, Qwen3MoEForCausalLMParameters as P,
                                                     Qwen3MoELayerParameters as L, Qwen3MoEParameters as B)
    else:
        from d9d_b200.module.model.mixtral import (MixtralForCausalLM as M, MixtralForCausalLMParameters as P,
                                                   MixtralLayerParameters as L, MixtralParameters as B)
    p = P(model=B(layer=L(hidden_size=32, intermediate_size=16, num_experts=4, experts_top_k=2, num_attention_heads=4,
                          num_key_value_heads=2, rms_norm_eps=1e-6, head_dim=8), num_hidden_layers=2, rope_base=10000,
                  max_position_ids=64, **VOCAB))
    m = M(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    m.reset_parameters()
    return p, m


@pytest.mark.parametrize("family", ["qwen3_moe", "mixtral"])
@pytest.mark.parametrize("fmt", ["module_list", "fused"])
def test_moe_round_trip(family, fmt):
    import importlib

    hfmod = importlib.import_module(f"d9d_b200.module.model.{family}.huggingface")
    p, m = _moe_model(family)
    native = {k: v.detach().clone() for k, v in m.state_dict().items()}
    to_hf = getattr(hfmod, f"mapper_to_huggingface_{family}_for_causal_lm")(p, fmt)
    from_hf = getattr(hfmod, f"mapper_from_huggingface_{family}_for_causal_lm")(p, fmt)
    hf_state = _run(to_hf, native)
    assert "lm_head.weight" in hf_state and "model.embed_tokens.weight" in hf_state
    if fmt == "module_list":
        key = "model.layers.0.mlp.experts.3.down_proj.weight" if family == "qwen3_moe" else "model.layers.1.block_sparse_moe.experts.3.w2.weight"
        assert hf_state[key].shape == (32, 16)  # HF nn.Linear layout [out, in]
    else:
        assert hf_state["model.layers.0.mlp.experts.gate_up_proj"].shape == (4, 32, 32)  # [E, 2I, H]
        assert hf_state["model.layers.0.mlp.experts.down_proj"].shape == (4, 32, 16)  # [E, H, I]
    back = _run(from_hf, hf_state)
    assert back.keys() == native.keys()
    for k in native:
        torch.testing.assert_close(back[k], native[k], rtol=0, atol=0)


@pytest.mark.parametrize("family", ["qwen3_dense", "llama3"])
@pytest.mark.parametrize("head", ["causal_lm", "classification", "embedding"])
def test_dense_round_trip(family, head):
    import importlib

    mod = importlib.import_module(f"d9d_b200.module.model.{family}")
    hfmod = importlib.import_module(f"d9d_b200.module.model.{family}.huggingface")
    cls = "Qwen3Dense" if family == "qwen3_dense" else "Llama3"
    base = getattr(mod, f"{cls}Parameters")(layer=getattr(mod, f"{cls}LayerParameters")(
        hidden_size=32, intermediate_size=48, num_attention_heads=4, num_key_value_heads=2, rms_norm_eps=1e-6, head_dim=8),
        num_hidden_layers=2, rope_base=10000, max_position_ids=64, **VOCAB)
    if head == "causal_lm":
        p = getattr(mod, f"{cls}ForCausalLMParameters")(model=base)
        m = getattr(mod, f"{cls}ForCausalLM")(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    elif head == "classification":
        p = getattr(mod, f"{cls}ForClassificationParameters")(model=base, num_labels=3, classifier_dropout=0.0)
        m = getattr(mod, f"{cls}ForClassification")(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.mean, False)
    else:
        p = getattr(mod, f"{cls}ForEmbeddingParameters")(model=base)
        m = getattr(mod, f"{cls}ForEmbedding")(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.mean, False)
    m.reset_parameters()
    native = {k: v.detach().clone() for k, v in m.state_dict().items()}
    hf_state = _run(getattr(hfmod, f"mapper_to_huggingface_{family}_for_{head}")(p), native)
    if head == "embedding":
        assert "embed_tokens.weight" in hf_state  # bare backbone
    back = _run(getattr(hfmod, f"mapper_from_huggingface_{family}_for_{head}")(p), hf_state)
    assert back.keys() == native.keys()
    for k in native:
        torch.testing.assert_close(back[k], native[k], rtol=0, atol=0)


def test_qwen3_dense_matches_transformers():
    transformers = pytest.importorskip("transformers")
    from d9d_b200.module.model.qwen3_dense import (Qwen3DenseForCausalLM, Qwen3DenseForCausalLMParameters,
                                                   Qwen3DenseLayerParameters, Qwen3DenseParameters)
    from d9d_b200.module.model.qwen3_dense.huggingface import mapper_from_huggingface_qwen3_dense_for_causal_lm

    cfg = transformers.Qwen3Config(vocab_size=96, hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=2, head_dim=8, rms_norm_eps=1e-6, rope_theta=10000.0, max_position_embeddings=64,
                                   tie_word_embeddings=False, attention_bias=False)
    torch.manual_seed(0)
    hf_model = transformers.Qwen3ForCausalLM(cfg).eval()
    p = Qwen3DenseForCausalLMParameters(model=Qwen3DenseParameters(
        layer=Qwen3DenseLayerParameters(hidden_size=32, intermediate_size=48, num_attention_heads=4, num_key_value_heads=2,
                                        rms_norm_eps=1e-6, head_dim=8), num_hidden_layers=2, rope_base=10000, max_position_ids=64, **VOCAB))
    ours = Qwen3DenseForCausalLM(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    ours.reset_parameters()
    state = _run(mapper_from_huggingface_qwen3_dense_for_causal_lm(p), dict(hf_model.state_dict()))
    missing, unexpected = ours.load_state_dict(state, strict=False)
    assert not unexpected and all("rope" in k or "cos" in k or "sin" in k for k in missing), (missing, unexpected)
    ids = torch.randint(0, 96, (2, 12))
    labels = torch.randint(0, 96, (2, 12))
    pos = torch.arange(12)[None].expand(2, -1)
    with torch.no_grad():
        logits = hf_model(input_ids=ids, position_ids=pos).logits.float()
        want = torch.nn.functional.cross_entropy(logits.view(-1, 96), labels.view(-1), reduction="none").view(2, 12)
        got = ours(input_ids=ids, position_ids=pos, labels=labels)["logps"]
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)


def test_qwen3_moe_matches_transformers():
    transformers = pytest.importorskip("transformers")
    from d9d_b200.module.model.qwen3_moe.huggingface import mapper_from_huggingface_qwen3_moe_for_causal_lm

    cfg = transformers.Qwen3MoeConfig(vocab_size=96, hidden_size=32, intermediate_size=48, moe_intermediate_size=16, num_hidden_layers=2,
                                      num_attention_heads=4, num_key_value_heads=2, head_dim=8, rms_norm_eps=1e-6, rope_theta=10000.0,
                                      max_position_embeddings=64, tie_word_embeddings=False, attention_bias=False, num_experts=4,
                                      num_experts_per_tok=2, norm_topk_prob=True, decoder_sparse_step=1, mlp_only_layers=[],
                                      router_aux_loss_coef=0.0, output_router_logits=False)
    torch.manual_seed(0)
    hf_model = transformers.Qwen3MoeForCausalLM(cfg).eval()
    hf_state = dict(hf_model.state_dict())
    fmt = "fused" if any(k.endswith("experts.gate_up_proj") for k in hf_state) else "module_list"
    p, ours = _moe_model("qwen3_moe")
    state = _run(mapper_from_huggingface_qwen3_moe_for_causal_lm(p, fmt), hf_state)
    missing, unexpected = ours.load_state_dict(state, strict=False)
    assert not unexpected and all("rope" in k or "cos" in k or "sin" in k or "tokens_per_expert" in k for k in missing), (missing, unexpected)
    ids = torch.randint(0, 96, (2, 12))
    labels = torch.randint(0, 96, (2, 12))
    pos = torch.arange(12)[None].expand(2, -1)
    with torch.no_grad():
        logits = hf_model(input_ids=ids, position_ids=pos).logits.float()
        want = torch.nn.functional.cross_entropy(logits.view(-1, 96), labels.view(-1), reduction="none").view(2, 12)
        got = ours(input_ids=ids, position_ids=pos, labels=labels)["logps"]
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)


def _load(ours, state):
    missing, unexpected = ours.load_state_dict(state, strict=False)
    assert not unexpected and all("rope" in k or "cos" in k or "sin" in k or "tokens_per_expert" in k for k in missing), (missing, unexpected)


def _per_token_nll(hf_model, ids, pos, labels, vocab=96):
    logits = hf_model(input_ids=ids, position_ids=pos).logits.float()
    return torch.nn.functional.cross_entropy(logits.view(-1, vocab), labels.view(-1), reduction="none").view(labels.shape)


def test_llama3_matches_transformers():
    transformers = pytest.importorskip("transformers")
    from d9d_b200.module.model.llama3 import (Llama3ForCausalLM, Llama3ForCausalLMParameters, Llama3LayerParameters, Llama3Parameters,
                                              mapper_from_huggingface_llama3_for_causal_lm)

    cfg = transformers.LlamaConfig(vocab_size=96, hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=2, head_dim=8, rms_norm_eps=1e-6, rope_theta=10000.0, max_position_embeddings=64,
                                   tie_word_embeddings=False, attention_bias=False, mlp_bias=False)
    torch.manual_seed(0)
    hf_model = transformers.LlamaForCausalLM(cfg).eval()
    p = Llama3ForCausalLMParameters(model=Llama3Parameters(
        layer=Llama3LayerParameters(hidden_size=32, intermediate_size=48, num_attention_heads=4, num_key_value_heads=2, rms_norm_eps=1e-6,
                                    head_dim=8), num_hidden_layers=2, rope_base=10000, max_position_ids=64, **VOCAB))
    ours = Llama3ForCausalLM(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    ours.reset_parameters()
    _load(ours, _run(mapper_from_huggingface_llama3_for_causal_lm(p), dict(hf_model.state_dict())))
    ids, labels = torch.randint(0, 96, (2, 12)), torch.randint(0, 96, (2, 12))
    pos = torch.arange(12)[None].expand(2, -1)
    with torch.no_grad():
        want = _per_token_nll(hf_model, ids, pos, labels)
        got = ours(input_ids=ids, position_ids=pos, labels=labels)["logps"]
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)


def test_mixtral_matches_transformers():
    transformers = pytest.importorskip("transformers")
    from d9d_b200.module.model.mixtral import mapper_from_huggingface_mixtral_for_causal_lm

    cfg = transformers.MixtralConfig(vocab_size=96, hidden_size=32, intermediate_size=16, num_hidden_layers=2, num_attention_heads=4,
                                     num_key_value_heads=2, head_dim=8, rms_norm_eps=1e-6, rope_theta=10000.0, max_position_embeddings=64,
                                     tie_word_embeddings=False, num_local_experts=4, num_experts_per_tok=2, router_aux_loss_coef=0.0,
                                     output_router_logits=False, sliding_window=None, router_jitter_noise=0.0)
    torch.manual_seed(0)
    hf_model = transformers.MixtralForCausalLM(cfg).eval()
    hf_state = dict(hf_model.state_dict())
    fmt = "fused" if any(k.endswith("experts.gate_up_proj") for k in hf_state) else "module_list"
    p, ours = _moe_model("mixtral")
    _load(ours, _run(mapper_from_huggingface_mixtral_for_causal_lm(p, fmt), hf_state))
    ids, labels = torch.randint(0, 96, (2, 12)), torch.randint(0, 96, (2, 12))
    pos = torch.arange(12)[None].expand(2, -1)
    with torch.no_grad():
        want = _per_token_nll(hf_model, ids, pos, labels)
        got = ours(input_ids=ids, position_ids=pos, labels=labels)["logps"]
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)


def test_qwen3_classification_and_embedding_match_transformers():
    """HF classifies on the last token of every row; we select the same tokens with ``pooling_mask``."""
    transformers = pytest.importorskip("transformers")
    from d9d_b200.module.model.qwen3_dense import (Qwen3DenseForClassification, Qwen3DenseForClassificationParameters,
                                                   Qwen3DenseForEmbedding, Qwen3DenseForEmbeddingParameters, Qwen3DenseLayerParameters,
                                                   Qwen3DenseParameters, mapper_from_huggingface_qwen3_dense_for_classification,
                                                   mapper_from_huggingface_qwen3_dense_for_embedding)

    cfg = transformers.Qwen3Config(vocab_size=96, hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=2, head_dim=8, rms_norm_eps=1e-6, rope_theta=10000.0, max_position_embeddings=64,
                                   tie_word_embeddings=False, attention_bias=False, num_labels=3, pad_token_id=None)
    torch.manual_seed(0)
    hf_cls = transformers.Qwen3ForSequenceClassification(cfg).eval()
    base = Qwen3DenseParameters(layer=Qwen3DenseLayerParameters(hidden_size=32, intermediate_size=48, num_attention_heads=4,
                                                                num_key_value_heads=2, rms_norm_eps=1e-6, head_dim=8),
                                num_hidden_layers=2, rope_base=10000, max_position_ids=64, **VOCAB)
    ids = torch.randint(0, 96, (1, 12))  # batch of one: without a pad token HF only accepts single rows
    pos = torch.arange(12)[None]
    last = torch.zeros(1, 12, dtype=torch.long)
    last[:, -1] = 1

    p = Qwen3DenseForClassificationParameters(model=base, num_labels=3, classifier_dropout=0.0)
    ours = Qwen3DenseForClassification(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False).eval()
    ours.reset_parameters()
    _load(ours, _run(mapper_from_huggingface_qwen3_dense_for_classification(p), dict(hf_cls.state_dict())))
    with torch.no_grad():
        want = hf_cls(input_ids=ids, position_ids=pos).logits.float()
        got = ours(input_ids=ids, position_ids=pos, pooling_mask=last)["scores"]
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)

    p = Qwen3DenseForEmbeddingParameters(model=base)
    ours = Qwen3DenseForEmbedding(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False).eval()
    ours.reset_parameters()
    _load(ours, _run(mapper_from_huggingface_qwen3_dense_for_embedding(p), dict(hf_cls.model.state_dict())))
    with torch.no_grad():
        hidden = hf_cls.model(input_ids=ids, position_ids=pos).last_hidden_state.float()
        got = ours(input_ids=ids, position_ids=pos, pooling_mask=last)["embeddings"]
    want = hidden[:, -1]
    if got.norm(dim=-1).sub(1).abs().max() < 1e-4:  # the head L2-normalises by default
        want = torch.nn.functional.normalize(want, dim=-1)
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("q_lora_rank", [None, 12])
def test_deepseek_v2_matches_transformers_and_round_trips(q_lora_rank):
    """Latent attention + dense first layer + MoE layers with shared experts, against ``DeepseekV2ForCausalLM``."""
    transformers = pytest.importorskip("transformers")
    from d9d_b200.module.model.deepseek_v2 import (DeepseekV2ForCausalLM, DeepseekV2ForCausalLMParameters, DeepseekV2LayerParameters,
                                                   DeepseekV2Parameters, mapper_from_huggingface_deepseek_v2_for_causal_lm,
                                                   mapper_to_huggingface_deepseek_v2_for_causal_lm)

    cfg = transformers.DeepseekV2Config(
        vocab_size=96, hidden_size=32, intermediate_size=48, moe_intermediate_size=16, num_hidden_layers=3, num_attention_heads=4,
        num_key_value_heads=4, first_k_dense_replace=1, kv_lora_rank=16, q_lora_rank=q_lora_rank, n_routed_experts=4,
        n_shared_experts=2, qk_nope_head_dim=8, qk_rope_head_dim=4, v_head_dim=8, num_experts_per_tok=2, topk_method="greedy",
        norm_topk_prob=False, routed_scaling_factor=1.0, rms_norm_eps=1e-6, rope_theta=10000.0, max_position_embeddings=64,
        tie_word_embeddings=False, attention_bias=False, mlp_bias=False)
    torch.manual_seed(0)
    hf_model = transformers.DeepseekV2ForCausalLM(cfg).eval()
    p = DeepseekV2ForCausalLMParameters(model=DeepseekV2Parameters(
        layer=DeepseekV2LayerParameters(hidden_size=32, rms_norm_eps=1e-6, num_attention_heads=4, qk_nope_head_dim=8, qk_rope_head_dim=4,
                                        v_head_dim=8, kv_lora_rank=16, q_lora_rank=q_lora_rank, intermediate_size=48,
                                        first_k_dense_replace=1, moe_intermediate_size=16, num_experts=4, experts_top_k=2,
                                        num_shared_experts=2),
        num_hidden_layers=3, rope_base=10000, max_position_ids=64, **VOCAB))
    ours = DeepseekV2ForCausalLM(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    ours.reset_parameters()
    from d9d_b200.module.block.ffn import SwiGLU
    from d9d_b200.module.block.moe import MoELayer

    assert isinstance(ours.model.layers["0"].mlp, SwiGLU) and isinstance(ours.model.layers["1"].mlp, MoELayer)

    hf_state = dict(hf_model.state_dict())
    fmt = "fused" if any(k.endswith("experts.gate_up_proj") for k in hf_state) else "module_list"
    _load(ours, _run(mapper_from_huggingface_deepseek_v2_for_causal_lm(p, fmt), hf_state))
    ids, labels = torch.randint(0, 96, (2, 12)), torch.randint(0, 96, (2, 12))
    pos = torch.arange(12)[None].expand(2, -1)
    with torch.no_grad():
        want = _per_token_nll(hf_model, ids, pos, labels)
        got = ours(input_ids=ids, position_ids=pos, labels=labels)["logps"]
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)

    native = {k: v.detach().clone() for k, v in ours.state_dict().items() if "tokens_per_expert" not in k}
    exported = _run(mapper_to_huggingface_deepseek_v2_for_causal_lm(p, fmt), native)
    assert exported.keys() == hf_state.keys()
    for k in hf_state:
        torch.testing.assert_close(exported[k], hf_state[k], rtol=0, atol=0)


def test_qwen3_5_hybrid_matches_transformers_and_round_trips():
    """Gated DeltaNet layers interleaved with gated partial-rotary attention, zero-centred norms, vs ``Qwen3_5ForCausalLM``."""
    pytest.importorskip("transformers")
    from transformers.models.qwen3_5.configuration_qwen3_5 import Qwen3_5TextConfig
    from transformers.models.qwen3_5.modeling_qwen3_5 import Qwen3_5ForCausalLM as HFModel

    from d9d_b200.module.block.attention import GatedDeltaNet, GroupedQueryAttention
    from d9d_b200.module.model.qwen3_5 import (Qwen3_5ForCausalLM, Qwen3_5ForCausalLMParameters, Qwen3_5LayerParameters, Qwen3_5Parameters,
                                               mapper_from_huggingface_qwen3_5_for_causal_lm, mapper_to_huggingface_qwen3_5_for_causal_lm)

    cfg = Qwen3_5TextConfig(vocab_size=96, hidden_size=32, intermediate_size=48, num_hidden_layers=4, num_attention_heads=4,
                            num_key_value_heads=2, head_dim=16, rms_norm_eps=1e-6, max_position_embeddings=64, tie_word_embeddings=False,
                            linear_conv_kernel_dim=4, linear_key_head_dim=8, linear_value_head_dim=8, linear_num_key_heads=2,
                            linear_num_value_heads=4, full_attention_interval=2,
                            rope_parameters={"rope_type": "default", "rope_theta": 10000.0, "partial_rotary_factor": 0.25})
    assert cfg.layer_types == ["linear_attention", "full_attention"] * 2
    torch.manual_seed(0)
    hf_model = HFModel(cfg).eval()
    with torch.no_grad():  # HF initialises the zero-centred norm weights at exactly 0: perturb them so that they matter
        for name, param in hf_model.named_parameters():
            if "norm" in name or "dt_bias" in name:
                param.add_(torch.randn_like(param) * 0.1)
    p = Qwen3_5ForCausalLMParameters(model=Qwen3_5Parameters(
        layer=Qwen3_5LayerParameters(hidden_size=32, intermediate_size=48, rms_norm_eps=1e-6, num_attention_heads=4, num_key_value_heads=2,
                                     head_dim=16, partial_rotary_factor=0.25, linear_num_key_heads=2, linear_num_value_heads=4,
                                     linear_key_head_dim=8, linear_value_head_dim=8, linear_conv_kernel_dim=4, full_attention_interval=2),
        num_hidden_layers=4, rope_base=10000, max_position_ids=64, **VOCAB))
    ours = Qwen3_5ForCausalLM(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    ours.reset_parameters()
    assert isinstance(ours.model.layers["0"].self_attn, GatedDeltaNet) and isinstance(ours.model.layers["1"].self_attn, GroupedQueryAttention)

    hf_state = dict(hf_model.state_dict())
    _load(ours, _run(mapper_from_huggingface_qwen3_5_for_causal_lm(p), hf_state))
    ids, labels = torch.randint(0, 96, (2, 12)), torch.randint(0, 96, (2, 12))
    pos = torch.arange(12)[None].expand(2, -1)
    with torch.no_grad():
        want = _per_token_nll(hf_model, ids, pos, labels)
        got = ours(input_ids=ids, position_ids=pos, labels=labels)["logps"]
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)

    exported = _run(mapper_to_huggingface_qwen3_5_for_causal_lm(p), {k: v.detach().clone() for k, v in ours.state_dict().items()})
    assert exported.keys() == hf_state.keys()
    for k in hf_state:
        torch.testing.assert_close(exported[k], hf_state[k], rtol=0, atol=0)


def test_deepseek_v3_routing_matches_transformers():
    """Sigmoid scores + selection bias + group-limited top-k + weight scaling, latent attention with a low-rank query, vs
    ``DeepseekV3ForCausalLM``."""
    transformers = pytest.importorskip("transformers")
    from d9d_b200.module.model.deepseek_v3 import (DeepseekV3ForCausalLM, DeepseekV3ForCausalLMParameters, DeepseekV3LayerParameters,
                                                   DeepseekV3Parameters, deepseek_v3_router, mapper_from_huggingface_deepseek_v3_for_causal_lm,
                                                   mapper_to_huggingface_deepseek_v3_for_causal_lm)

    cfg = transformers.DeepseekV3Config(
        vocab_size=96, hidden_size=32, intermediate_size=48, moe_intermediate_size=16, num_hidden_layers=3, num_attention_heads=4,
        num_key_value_heads=4, first_k_dense_replace=1, kv_lora_rank=16, q_lora_rank=12, n_routed_experts=8, n_shared_experts=1,
        qk_nope_head_dim=8, qk_rope_head_dim=4, v_head_dim=8, num_experts_per_tok=3, n_group=4, topk_group=2, norm_topk_prob=True,
        routed_scaling_factor=2.5, rms_norm_eps=1e-6, rope_theta=10000.0, max_position_embeddings=64, tie_word_embeddings=False,
        attention_bias=False)
    torch.manual_seed(0)
    hf_model = transformers.DeepseekV3ForCausalLM(cfg).eval()
    with torch.no_grad():
        for name, buf in hf_model.named_buffers():
            if name.endswith("e_score_correction_bias"):
                buf.copy_(torch.randn_like(buf) * 0.05)  # make the selection bias matter
    p = DeepseekV3ForCausalLMParameters(model=DeepseekV3Parameters(
        layer=DeepseekV3LayerParameters(hidden_size=32, rms_norm_eps=1e-6, num_attention_heads=4, qk_nope_head_dim=8, qk_rope_head_dim=4,
                                        v_head_dim=8, kv_lora_rank=16, q_lora_rank=12, intermediate_size=48, first_k_dense_replace=1,
                                        moe_intermediate_size=16, num_experts=8, experts_top_k=3, num_shared_experts=1,
                                        router_renormalize_probabilities=True,
                                        router=deepseek_v3_router(n_group=4, topk_group=2, routed_scaling_factor=2.5)),
        num_hidden_layers=3, rope_base=10000, max_position_ids=64, **VOCAB))
    ours = DeepseekV3ForCausalLM(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    ours.reset_parameters()
    hf_state = dict(hf_model.state_dict())
    fmt = "fused" if any(k.endswith("experts.gate_up_proj") for k in hf_state) else "module_list"
    _load(ours, _run(mapper_from_huggingface_deepseek_v3_for_causal_lm(p, fmt), hf_state))
    ids, labels = torch.randint(0, 96, (2, 12)), torch.randint(0, 96, (2, 12))
    pos = torch.arange(12)[None].expand(2, -1)
    with torch.no_grad():
        want = _per_token_nll(hf_model, ids, pos, labels)
        got = ours(input_ids=ids, position_ids=pos, labels=labels)["logps"]
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)
    exported = _run(mapper_to_huggingface_deepseek_v3_for_causal_lm(p, fmt), {k: v.detach().clone() for k, v in ours.state_dict().items()})
    assert exported.keys() == hf_state.keys()


def test_qwen3_5_moe_matches_transformers_and_round_trips():
    """Hybrid token mixers + MoE feed-forward with a sigmoid-gated shared expert, vs ``Qwen3_5MoeForCausalLM``."""
    pytest.importorskip("transformers")
    from transformers.models.qwen3_5_moe.configuration_qwen3_5_moe import Qwen3_5MoeTextConfig
    from transformers.models.qwen3_5_moe.modeling_qwen3_5_moe import Qwen3_5MoeForCausalLM as HFModel

    from d9d_b200.module.model.qwen3_5_moe import (Qwen3_5MoEForCausalLM, Qwen3_5MoEForCausalLMParameters, Qwen3_5MoELayerParameters,
                                                   Qwen3_5MoEParameters, mapper_from_huggingface_qwen3_5_moe_for_causal_lm,
                                                   mapper_to_huggingface_qwen3_5_moe_for_causal_lm)

    cfg = Qwen3_5MoeTextConfig(vocab_size=96, hidden_size=32, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16,
                               rms_norm_eps=1e-6, max_position_embeddings=64, tie_word_embeddings=False, linear_conv_kernel_dim=4,
                               linear_key_head_dim=8, linear_value_head_dim=8, linear_num_key_heads=2, linear_num_value_heads=4,
                               moe_intermediate_size=16, shared_expert_intermediate_size=24, num_experts_per_tok=2, num_experts=4,
                               full_attention_interval=2,
                               rope_parameters={"rope_type": "default", "rope_theta": 10000.0, "partial_rotary_factor": 0.25})
    torch.manual_seed(0)
    hf_model = HFModel(cfg).eval()
    with torch.no_grad():
        for name, param in hf_model.named_parameters():
            if "norm" in name or "dt_bias" in name:
                param.add_(torch.randn_like(param) * 0.1)
            if name.endswith("mlp.gate.weight") or "experts" in name or "shared_expert_gate" in name:
                param.normal_(0, 0.2)  # HF initialises the router to zeros
    p = Qwen3_5MoEForCausalLMParameters(model=Qwen3_5MoEParameters(
        layer=Qwen3_5MoELayerParameters(hidden_size=32, rms_norm_eps=1e-6, num_attention_heads=4, num_key_value_heads=2, head_dim=16,
                                        partial_rotary_factor=0.25, linear_num_key_heads=2, linear_num_value_heads=4,
                                        linear_key_head_dim=8, linear_value_head_dim=8, linear_conv_kernel_dim=4, full_attention_interval=2,
                                        moe_intermediate_size=16, shared_expert_intermediate_size=24, num_experts=4, experts_top_k=2),
        num_hidden_layers=4, rope_base=10000, max_position_ids=64, **VOCAB))
    ours = Qwen3_5MoEForCausalLM(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    ours.reset_parameters()
    hf_state = dict(hf_model.state_dict())
    fmt = "fused" if any(k.endswith("experts.gate_up_proj") for k in hf_state) else "module_list"
    _load(ours, _run(mapper_from_huggingface_qwen3_5_moe_for_causal_lm(p, fmt), hf_state))
    ids, labels = torch.randint(0, 96, (2, 12)), torch.randint(0, 96, (2, 12))
    pos = torch.arange(12)[None].expand(2, -1)
    with torch.no_grad():
        want = _per_token_nll(hf_model, ids, pos, labels)
        got = ours(input_ids=ids, position_ids=pos, labels=labels)["logps"]
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)
    exported = _run(mapper_to_huggingface_qwen3_5_moe_for_causal_lm(p, fmt), {k: v.detach().clone() for k, v in ours.state_dict().items()})
    assert exported.keys() == hf_state.keys()
    for k in hf_state:
        torch.testing.assert_close(exported[k], hf_state[k], rtol=0, atol=0)


def test_mapper_groups_whose_inputs_live_in_different_shard_files(tmp_path):
    """Streaming load keeps a group's early inputs resident until the shard holding the rest has been read."""
    from tests.test_huggingface_mappers import _moe_model
    from d9d_b200.model_state.io import load_model_state, save_model_state
    from d9d_b200.module.model.qwen3_moe import mapper_from_huggingface_qwen3_moe_for_causal_lm, mapper_to_huggingface_qwen3_moe_for_causal_lm
    p, m = _moe_model("qwen3_moe")
    # export in the per-expert layout with shards so small that the experts of one layer land in many files
    save_model_state(tmp_path / "hf", mapper_to_huggingface_qwen3_moe_for_causal_lm(p, "module_list"), m, shard_size_gb=1.3e-5, show_progress=False)
    index = json.loads((tmp_path / "hf" / "model.safetensors.index.json").read_text())
    files = {index["weight_map"][f"model.layers.0.mlp.experts.{e}.up_proj.weight"] for e in range(4)}
    assert len(files) > 1
    _, clone = _moe_model("qwen3_moe")
    with torch.no_grad():
        for prm in clone.parameters(): prm.zero_()
    load_model_state(tmp_path / "hf", mapper_from_huggingface_qwen3_moe_for_causal_lm(p, "module_list"), "cpu", clone, show_progress=False)
    for k, v in m.state_dict().items():
        assert torch.equal(clone.state_dict()[k], v), k
