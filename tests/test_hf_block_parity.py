"""Block-level numerical parity with the HuggingFace ``transformers`` implementations (CPU, fp32, tiny shapes).

Mirrors the reference's block tests (``test/d9d_test/modules/block/**/test_hf_*.py``): weights are copied from the HF
module into ours, outputs are compared directly, gradients by angle + norm.
"""

import math

import pytest
import torch

transformers = pytest.importorskip("transformers")

from d9d_b200.module.block.positional import RotaryEmbeddingStyle  # noqa: E402


def _assert_grads_close(ours: torch.Tensor, theirs: torch.Tensor, name: str = "") -> None:
    a, b = ours.flatten().double(), theirs.flatten().double()
    cos = torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)
    assert cos > 1 - 1e-5, f"{name}: gradient direction differs (cos={cos})"
    assert abs(a.norm() / (b.norm() + 1e-30) - 1) < 1e-3, f"{name}: gradient norm differs"


def _copy(dst: torch.Tensor, src: torch.Tensor) -> None:
    with torch.no_grad():
        assert dst.shape == src.shape, (dst.shape, src.shape)
        dst.copy_(src)


def _cos_sin(head_dim: int, seq: int, base: float = 10000.0):
    inv = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    ang = torch.outer(torch.arange(seq).float(), inv)
    emb = torch.cat([ang, ang], dim=-1)
    return emb.cos()[None], emb.sin()[None]


def _causal_mask(seq: int) -> torch.Tensor:
    m = torch.full((seq, seq), float("-inf")).triu(1)
    return m[None, None]


# ---------------------------------------------------------------------------------------------- gated delta net
def test_gated_deltanet_matches_qwen3_5():
    from transformers.models.qwen3_5.configuration_qwen3_5 import Qwen3_5TextConfig
    from transformers.models.qwen3_5.modeling_qwen3_5 import Qwen3_5GatedDeltaNet

    from d9d_b200.module.block.attention.linear.gated_deltanet import GatedDeltaNet, MambaDecayGateParameters

    torch.manual_seed(0)
    cfg = Qwen3_5TextConfig(hidden_size=48, linear_num_key_heads=2, linear_num_value_heads=4, linear_key_head_dim=8,
                            linear_value_head_dim=16, linear_conv_kernel_dim=4, rms_norm_eps=1e-6, num_hidden_layers=1)
    hf = Qwen3_5GatedDeltaNet(cfg, layer_idx=0).float()
    with torch.no_grad():
        hf.dt_bias.uniform_(-1, 1)
        hf.norm.weight.uniform_(0.5, 1.5)
    ours = GatedDeltaNet(hidden_size=48, num_query_key_heads=2, num_value_heads=4, head_qk_dim=8, head_v_dim=16, norm_eps=1e-6,
                         conv_size=4, decay_gate=MambaDecayGateParameters(normalizer=16.0, dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4))
    _copy(ours.qkv_proj.weight, hf.in_proj_qkv.weight)
    _copy(ours.g_proj.weight, hf.in_proj_z.weight)
    _copy(ours.b_proj.weight, hf.in_proj_b.weight)
    _copy(ours.decay_gate.proj.weight, hf.in_proj_a.weight)
    _copy(ours.decay_gate.A_log, hf.A_log)
    _copy(ours.decay_gate.dt_bias, hf.dt_bias)
    _copy(ours.qkv_conv1d.weight, hf.conv1d.weight.squeeze(1))
    _copy(ours.out_norm.weight, hf.norm.weight)
    _copy(ours.o_proj.weight, hf.out_proj.weight)

    x = torch.randn(2, 70, 48)  # 70: one full 64-token chunk and a ragged tail
    with torch.no_grad():  # tiny activations (freshly initialised models): the q/k l2-norm's epsilon convention matters here
        small = x * 0.02
        torch.testing.assert_close(ours(small), hf(small), atol=1e-7, rtol=1e-3)
    x1, x2 = x.clone().requires_grad_(), x.clone().requires_grad_()
    y_hf = hf(x1)
    y = ours(x2)
    torch.testing.assert_close(y, y_hf, atol=2e-5, rtol=1e-4)
    w = torch.randn_like(y)
    (y_hf * w).sum().backward()
    (y * w).sum().backward()
    _assert_grads_close(x2.grad, x1.grad, "input")
    _assert_grads_close(ours.qkv_proj.weight.grad, hf.in_proj_qkv.weight.grad, "qkv")
    _assert_grads_close(ours.decay_gate.A_log.grad, hf.A_log.grad, "A_log")
    _assert_grads_close(ours.decay_gate.dt_bias.grad, hf.dt_bias.grad, "dt_bias")
    _assert_grads_close(ours.qkv_conv1d.weight.grad, hf.conv1d.weight.grad.squeeze(1), "conv")
    _assert_grads_close(ours.out_norm.weight.grad, hf.norm.weight.grad, "norm")


# ------------------------------------------------------------------------------ gated attention, partial rotary
def test_gated_partial_rope_attention_matches_qwen3_5():
    from transformers.models.qwen3_5.configuration_qwen3_5 import Qwen3_5TextConfig
    from transformers.models.qwen3_5.modeling_qwen3_5 import Qwen3_5Attention

    from d9d_b200.module.block.attention import GroupedQueryAttention

    torch.manual_seed(1)
    heads, kv_heads, dim, hidden, seq, rope_dim = 4, 2, 16, 40, 12, 8
    cfg = Qwen3_5TextConfig(hidden_size=hidden, num_attention_heads=heads, num_key_value_heads=kv_heads, head_dim=dim,
                            rms_norm_eps=1e-6, num_hidden_layers=1)
    cfg._attn_implementation = "eager"
    hf = Qwen3_5Attention(cfg, layer_idx=0).float()
    with torch.no_grad():
        hf.q_norm.weight.uniform_(-0.3, 0.3)  # zero-centred weights: scale is (1 + w)
        hf.k_norm.weight.uniform_(-0.3, 0.3)
    ours = GroupedQueryAttention(hidden_size=hidden, num_attention_heads=heads, num_key_value_heads=kv_heads, head_dim=dim,
                                 qk_norm_eps=1e-6, is_causal=True, rope_style=RotaryEmbeddingStyle.HALF, rope_dim=rope_dim,
                                 enable_output_gate=True, qk_norm_zero_centered=True)
    fused = hf.q_proj.weight.view(heads, 2 * dim, hidden)  # HF interleaves [query | gate] per head
    _copy(ours.q_proj.weight, fused[:, :dim].reshape(heads * dim, hidden))
    _copy(ours.gate_proj.weight, fused[:, dim:].reshape(heads * dim, hidden))
    for name in ("k_proj", "v_proj", "o_proj", "q_norm", "k_norm"):
        _copy(getattr(ours, name).weight, getattr(hf, name).weight)

    x = torch.randn(2, seq, hidden)
    cos, sin = _cos_sin(rope_dim, seq)  # partial rotary: only the first ``rope_dim`` dims of every head rotate
    y_hf, _ = hf(x, position_embeddings=(cos, sin), attention_mask=_causal_mask(seq))
    y = ours(x, None, (cos.expand(2, -1, -1), sin.expand(2, -1, -1)))
    torch.testing.assert_close(y, y_hf, atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("qk_norm", [True, False])
def test_grouped_query_attention_matches_qwen3_and_llama(qk_norm):
    from d9d_b200.module.block.attention import GroupedQueryAttention

    torch.manual_seed(2)
    heads, kv_heads, dim, hidden, seq = 4, 1, 8, 24, 10
    if qk_norm:
        from transformers.models.qwen3.configuration_qwen3 import Qwen3Config as Config
        from transformers.models.qwen3.modeling_qwen3 import Qwen3Attention as Attention
    else:
        from transformers.models.llama.configuration_llama import LlamaConfig as Config
        from transformers.models.llama.modeling_llama import LlamaAttention as Attention
    cfg = Config(hidden_size=hidden, num_attention_heads=heads, num_key_value_heads=kv_heads, head_dim=dim, rms_norm_eps=1e-6,
                 num_hidden_layers=1)
    cfg._attn_implementation = "eager"
    hf = Attention(cfg, layer_idx=0).float()
    ours = GroupedQueryAttention(hidden_size=hidden, num_attention_heads=heads, num_key_value_heads=kv_heads, head_dim=dim,
                                 qk_norm_eps=1e-6 if qk_norm else None, is_causal=True, rope_style=RotaryEmbeddingStyle.HALF)
    names = ["q_proj", "k_proj", "v_proj", "o_proj"] + (["q_norm", "k_norm"] if qk_norm else [])
    if qk_norm:
        with torch.no_grad():
            hf.q_norm.weight.uniform_(0.5, 1.5)
            hf.k_norm.weight.uniform_(0.5, 1.5)
    for name in names:
        _copy(getattr(ours, name).weight, getattr(hf, name).weight)
    x = torch.randn(3, seq, hidden)
    x1, x2 = x.clone().requires_grad_(), x.clone().requires_grad_()
    cos, sin = _cos_sin(dim, seq)
    y_hf, _ = hf(x1, position_embeddings=(cos, sin), attention_mask=_causal_mask(seq))
    y = ours(x2, None, (cos.expand(3, -1, -1), sin.expand(3, -1, -1)))
    torch.testing.assert_close(y, y_hf, atol=2e-5, rtol=1e-4)
    y_hf.square().sum().backward()
    y.square().sum().backward()
    _assert_grads_close(x2.grad, x1.grad, "input")
    for name in names:
        _assert_grads_close(getattr(ours, name).weight.grad, getattr(hf, name).weight.grad, name)


# ------------------------------------------------------------------------------------------ multi-head latent
@pytest.mark.parametrize("q_lora_rank", [None, 12])
def test_multi_head_latent_attention_matches_deepseek_v3(q_lora_rank):
    from transformers.models.deepseek_v3.configuration_deepseek_v3 import DeepseekV3Config
    from transformers.models.deepseek_v3.modeling_deepseek_v3 import DeepseekV3Attention

    from d9d_b200.module.block.attention import MultiHeadLatentAttention

    torch.manual_seed(3)
    hidden, heads, nope, rope, vdim, kv_rank, seq = 32, 4, 8, 4, 8, 16, 9
    cfg = DeepseekV3Config(hidden_size=hidden, num_attention_heads=heads, num_key_value_heads=heads, qk_nope_head_dim=nope,
                           qk_rope_head_dim=rope, v_head_dim=vdim, kv_lora_rank=kv_rank, q_lora_rank=q_lora_rank,
                           rms_norm_eps=1e-6, num_hidden_layers=1, attention_bias=False)
    cfg._attn_implementation = "eager"
    hf = DeepseekV3Attention(cfg, layer_idx=0).float()
    ours = MultiHeadLatentAttention(hidden_size=hidden, num_attention_heads=heads, qk_nope_head_dim=nope, qk_rope_head_dim=rope,
                                    v_head_dim=vdim, kv_lora_rank=kv_rank, q_lora_rank=q_lora_rank, qk_down_norm_eps=1e-6,
                                    is_causal=True, rope_style=RotaryEmbeddingStyle.INTERLEAVED)
    with torch.no_grad():
        hf.kv_a_layernorm.weight.uniform_(0.5, 1.5)
    if q_lora_rank is None:
        _copy(ours.q_proj.weight, hf.q_proj.weight)
    else:
        with torch.no_grad():
            hf.q_a_layernorm.weight.uniform_(0.5, 1.5)
        _copy(ours.q_proj.down_proj.weight, hf.q_a_proj.weight)
        _copy(ours.q_proj.norm.weight, hf.q_a_layernorm.weight)
        _copy(ours.q_proj.up_proj.weight, hf.q_b_proj.weight)
    _copy(ours.kv_down_proj.weight, hf.kv_a_proj_with_mqa.weight)
    _copy(ours.kv_down_norm.weight, hf.kv_a_layernorm.weight)
    _copy(ours.kv_up_proj.weight, hf.kv_b_proj.weight)
    _copy(ours.o_proj.weight, hf.o_proj.weight)

    x = torch.randn(2, seq, hidden)
    inv = 1.0 / (10000.0 ** (torch.arange(0, rope, 2).float() / rope))
    ang = torch.outer(torch.arange(seq).float(), inv)
    hf_emb = torch.cat([ang, ang], dim=-1)  # HF de-interleaves q/k and then rotates halves
    our_emb = ang.repeat_interleave(2, dim=-1)  # we rotate the interleaved pairs in place
    y_hf, _ = hf(x, position_embeddings=(hf_emb.cos()[None], hf_emb.sin()[None]), attention_mask=_causal_mask(seq))
    y = ours(x, None, (our_emb.cos()[None].expand(2, -1, -1), our_emb.sin()[None].expand(2, -1, -1)))
    torch.testing.assert_close(y, y_hf, atol=2e-5, rtol=1e-4)


# --------------------------------------------------------------------------------------------------- MoE block
@pytest.mark.parametrize("renormalize", [True, False])
def test_moe_layer_matches_qwen3_moe_block(renormalize):
    from transformers.models.qwen3_moe.configuration_qwen3_moe import Qwen3MoeConfig
    from transformers.models.qwen3_moe.modeling_qwen3_moe import Qwen3MoeSparseMoeBlock

    from d9d_b200.module.block.moe import MoELayer

    torch.manual_seed(4)
    hidden, inter, experts, top_k = 24, 12, 6, 3
    cfg = Qwen3MoeConfig(hidden_size=hidden, moe_intermediate_size=inter, num_experts=experts, num_experts_per_tok=top_k,
                         norm_topk_prob=renormalize, num_hidden_layers=1)
    hf = Qwen3MoeSparseMoeBlock(cfg).float()
    with torch.no_grad():
        for p in hf.parameters():
            p.normal_(0, 0.2)
    ours = MoELayer(hidden_dim=hidden, intermediate_dim_grouped=inter, num_grouped_experts=experts, top_k=top_k,
                    router_renormalize_probabilities=renormalize)
    ours.reset_parameters()
    _copy(ours.router.gate.weight, hf.gate.weight)
    gate_up = hf.experts.gate_up_proj  # [E, 2I, H]
    _copy(ours.grouped_experts.gate_proj.weight, gate_up[:, :inter].transpose(1, 2))
    _copy(ours.grouped_experts.up_proj.weight, gate_up[:, inter:].transpose(1, 2))
    _copy(ours.grouped_experts.down_proj.weight, hf.experts.down_proj.transpose(1, 2))  # [E, H, I] -> [E, I, H]

    x = torch.randn(2, 7, hidden)
    x1, x2 = x.clone().requires_grad_(), x.clone().requires_grad_()
    y_hf = hf(x1)
    y_hf = y_hf[0] if isinstance(y_hf, tuple) else y_hf
    y = ours(x2)
    torch.testing.assert_close(y, y_hf.view_as(y), atol=2e-5, rtol=1e-4)
    assert int(ours.tokens_per_expert.sum()) == 2 * 7 * top_k
    y_hf.square().sum().backward()
    y.square().sum().backward()
    _assert_grads_close(x2.grad, x1.grad, "input")
    _assert_grads_close(ours.router.gate.weight.grad, hf.gate.weight.grad, "router")
    _assert_grads_close(ours.grouped_experts.down_proj.weight.grad, hf.experts.down_proj.grad.transpose(1, 2), "down")


def test_rope_tables_match_transformers_scalings():
    """Linear / dynamic-NTK-free / YaRN inverse frequencies against ``transformers.modeling_rope_utils``."""
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    from transformers.models.llama.configuration_llama import LlamaConfig

    from d9d_b200.module.block.positional.rope_scaling import LinearRopeScaling, YarnRopeScaling

    dim, base = 32, 10000
    cfg = LlamaConfig(hidden_size=dim * 2, num_attention_heads=2, head_dim=dim, max_position_embeddings=128,
                      rope_parameters={"rope_type": "linear", "factor": 4.0, "rope_theta": base})
    inv, _ = ROPE_INIT_FUNCTIONS["linear"](cfg, device="cpu")
    torch.testing.assert_close(LinearRopeScaling(4.0).inverse_frequencies(base, dim).float(), inv.float())

    cfg = LlamaConfig(hidden_size=dim * 2, num_attention_heads=2, head_dim=dim, max_position_embeddings=512,
                      rope_parameters={"rope_type": "yarn", "factor": 8.0, "beta_fast": 32.0, "beta_slow": 1.0, "rope_theta": base,
                                       "original_max_position_embeddings": 64, "truncate": False})  # un-truncated ramp bounds, like the reference
    inv, mscale = ROPE_INIT_FUNCTIONS["yarn"](cfg, device="cpu")
    yarn = YarnRopeScaling(factor=8.0, beta_fast=32.0, beta_slow=1.0, original_max_position_embeddings=64)
    torch.testing.assert_close(yarn.inverse_frequencies(base, dim).float(), inv.float(), atol=1e-7, rtol=1e-5)
    assert math.isclose(yarn.attention_mscale, mscale, rel_tol=1e-6)
