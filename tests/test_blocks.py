"""Block-level behaviour on CPU (the fp32 PyTorch paths that also serve as oracles for the CUDA kernels)."""

import math

import pytest
import torch


def _rope(head_dim, max_pos=64, style=None, scaling=None):
    from d9d_b200.module.block.positional import RotaryEmbeddingProvider, RotaryEmbeddingStyle

    p = RotaryEmbeddingProvider(rope_base=10000, head_dim=head_dim, max_position_ids=max_pos,
                                style=style or RotaryEmbeddingStyle.HALF, rope_scaling=scaling)
    p.reset_parameters()
    return p


def test_rope_is_a_rotation_and_encodes_relative_positions():
    from d9d_b200.module.block.positional import RotaryEmbeddingApplicator, RotaryEmbeddingStyle

    for style in (RotaryEmbeddingStyle.HALF, RotaryEmbeddingStyle.INTERLEAVED):
        prov = _rope(16, style=style)
        app = RotaryEmbeddingApplicator(style)
        q, k = torch.randn(1, 8, 2, 16), torch.randn(1, 8, 2, 16)
        pos = torch.arange(8)[None]
        cos, sin = prov(pos)
        q1, k1 = app(q, k, cos, sin)
        torch.testing.assert_close(q1.norm(dim=-1), q.norm(dim=-1), rtol=1e-5, atol=1e-5)  # rotations preserve norms
        cos2, sin2 = prov(pos + 5)  # shifting every position leaves q·k unchanged
        q2, k2 = app(q, k, cos2, sin2)
        torch.testing.assert_close(torch.einsum("bshd,bthd->bhst", q1, k1), torch.einsum("bshd,bthd->bhst", q2, k2), rtol=1e-4, atol=1e-4)


def test_rope_scalings_change_frequencies_as_specified():
    from d9d_b200.module.block.positional.rope_scaling import LinearRopeScaling, NoRopeScaling, NtkRopeScaling, YarnRopeScaling

    base = NoRopeScaling().inverse_frequencies(10000, 32)
    torch.testing.assert_close(LinearRopeScaling(4.0).inverse_frequencies(10000, 32), base / 4.0)
    ntk = NtkRopeScaling(4.0).inverse_frequencies(10000, 32)
    assert abs(float(ntk[0] / base[0]) - 1.0) < 1e-6 and float(ntk[-1] / base[-1]) < 0.3  # high freq kept, low stretched
    yarn = YarnRopeScaling(factor=4.0, beta_fast=32.0, beta_slow=1.0, original_max_position_embeddings=64)
    f = yarn.inverse_frequencies(10000, 32)
    assert torch.all(f <= base + 1e-9) and torch.all(f >= base / 4.0 - 1e-9) and yarn.attention_mscale > 1.0
    with pytest.raises(ValueError):
        YarnRopeScaling(4.0, 1.0, 32.0, 64)


def test_gqa_equals_explicit_attention_and_is_causal():
    from d9d_b200.module.block.attention import GroupedQueryAttention
    from d9d_b200.module.block.positional import RotaryEmbeddingStyle

    torch.manual_seed(0)
    attn = GroupedQueryAttention(hidden_size=32, num_attention_heads=4, num_key_value_heads=2, head_dim=8, qk_norm_eps=1e-6,
                                 is_causal=True, rope_style=RotaryEmbeddingStyle.HALF, enable_output_gate=True)
    attn.reset_parameters()
    x = torch.randn(2, 10, 32)
    pe = _rope(8)(torch.arange(10)[None].expand(2, -1))
    y = attn(x, None, pe)
    x2 = x.clone()
    x2[:, 6:] += 1.0
    torch.testing.assert_close(attn(x2, None, pe)[:, :6], y[:, :6], rtol=1e-5, atol=1e-5)
    # explicit recomputation
    q = attn.q_norm(attn.q_proj(x).view(2, 10, 4, 8))
    k = attn.k_norm(attn.k_proj(x).view(2, 10, 2, 8))
    v = attn.v_proj(x).view(2, 10, 2, 8)
    q, k = attn.rope(q, k, *pe)
    k, v = k.repeat_interleave(2, dim=2), v.repeat_interleave(2, dim=2)
    s = torch.einsum("bshd,bthd->bhst", q, k) / math.sqrt(8)
    s = s.masked_fill(torch.triu(torch.ones(10, 10, dtype=torch.bool), 1), float("-inf"))
    o = torch.einsum("bhst,bthd->bshd", s.softmax(-1), v).reshape(2, 10, 32)
    want = attn.o_proj(o * torch.sigmoid(attn.gate_proj(x)))
    torch.testing.assert_close(y, want, rtol=1e-4, atol=1e-5)


def test_mla_shapes_causality_and_gradients():
    from d9d_b200.module.block.attention import MultiHeadLatentAttention
    from d9d_b200.module.block.positional import RotaryEmbeddingStyle

    torch.manual_seed(0)
    mla = MultiHeadLatentAttention(hidden_size=32, num_attention_heads=4, qk_nope_head_dim=8, qk_rope_head_dim=4, v_head_dim=6,
                                   kv_lora_rank=16, q_lora_rank=12, qk_down_norm_eps=1e-6, is_causal=True,
                                   rope_style=RotaryEmbeddingStyle.INTERLEAVED)
    mla.reset_parameters()
    x = torch.randn(2, 9, 32, requires_grad=True)
    pe = _rope(4, style=RotaryEmbeddingStyle.INTERLEAVED)(torch.arange(9)[None].expand(2, -1))
    y = mla(x, None, pe)
    assert y.shape == (2, 9, 32)
    y.square().sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in mla.parameters())
    x2 = x.detach().clone()
    x2[:, 5:] -= 2.0
    torch.testing.assert_close(mla(x2, None, pe)[:, :5], y[:, :5].detach(), rtol=1e-4, atol=1e-5)


def test_split_embeddings_and_lm_head_agree_with_a_single_vocab():
    from d9d_b200.module.block.embedding import SplitTokenEmbeddings
    from d9d_b200.module.block.head import LM_IGNORE_INDEX, SplitLanguageModellingHead

    torch.manual_seed(0)
    sizes, order = {"regular": 20, "special": 5}, ["regular", "special"]
    emb = SplitTokenEmbeddings(sizes, order, 16)
    emb.reset_parameters()
    ids = torch.tensor([[0, 19, 20, 24]])
    table = torch.cat([emb.token_embedding["regular"].weight, emb.token_embedding["special"].weight])
    torch.testing.assert_close(emb(ids), table[ids])

    head = SplitLanguageModellingHead(sizes, order, 16)
    head.reset_parameters()
    h = torch.randn(2, 6, 16)
    labels = torch.randint(0, 25, (2, 6))
    labels[0, 0] = LM_IGNORE_INDEX
    w = torch.cat([head.lm_head["regular"].weight, head.lm_head["special"].weight])
    want = torch.nn.functional.cross_entropy((h @ w.t()).view(-1, 25), labels.view(-1), reduction="none", ignore_index=LM_IGNORE_INDEX).view(2, 6)
    got = head(hidden_states=h, labels=labels)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    assert float(got[0, 0]) == 0.0 and got.dtype == torch.float32


def test_pooled_heads():
    from d9d_b200.module.block.head import ClassificationHead, EmbeddingHead

    torch.manual_seed(0)
    h = torch.randn(2, 5, 8)
    mask = torch.tensor([[0, 0, 1, 0, 0], [1, 1, 0, 0, 0]])
    cls = ClassificationHead(hidden_size=8, num_labels=3, dropout=0.0)
    cls.reset_parameters()
    scores = cls(hidden_states=h, pooling_mask=mask)
    want = h[mask == 1] @ cls.score.weight.t()  # one row per selected token (mean pooling is the aggregator's job)
    torch.testing.assert_close(scores, want.float(), rtol=1e-5, atol=1e-5)
    eh = EmbeddingHead(hidden_size=8, embedding_dim=4, normalize=True)
    eh.reset_parameters()
    e = eh(hidden_states=h, pooling_mask=mask)
    assert e.shape == (3, 4)
    torch.testing.assert_close(e.norm(dim=-1), torch.ones(3), rtol=1e-5, atol=1e-5)


def test_moe_layer_matches_per_token_expert_sum_and_counts_tokens():
    from d9d_b200.module.block.moe import MoELayer

    torch.manual_seed(0)
    moe = MoELayer(hidden_dim=16, intermediate_dim_grouped=8, num_grouped_experts=4, top_k=2, router_renormalize_probabilities=True)
    moe.reset_parameters()
    moe.reset_stats()
    x = torch.randn(3, 7, 16)
    y = moe(x)
    flat = x.view(-1, 16)
    probs = (flat @ moe.router.gate.weight.t()).float().softmax(-1)
    top_p, top_i = probs.topk(2, dim=-1)
    top_p = top_p / top_p.sum(-1, keepdim=True)
    want = torch.zeros_like(flat)
    ge = moe.grouped_experts
    for t in range(flat.shape[0]):
        for p, e in zip(top_p[t], top_i[t]):
            gate, up = flat[t] @ ge.gate_proj.weight[e], flat[t] @ ge.up_proj.weight[e]
            want[t] += p * ((torch.nn.functional.silu(gate) * up) @ ge.down_proj.weight[e])
    torch.testing.assert_close(y.view(-1, 16), want, rtol=1e-4, atol=1e-5)
    assert int(moe.tokens_per_expert.sum()) == 21 * 2


def test_reference_compatible_moe_routing_helpers():
    from d9d_b200.kernel.moe import fused_indices_to_multihot, moe_permute_with_probs, moe_unpermute_mask

    idx = torch.tensor([[0, 2], [1, -1], [3, 0]])
    p = torch.tensor([[0.6, 0.4], [1.0, 0.5], [0.3, 0.7]])
    routing, probs = fused_indices_to_multihot(idx, p, 4)
    assert routing.tolist() == [[True, False, True, False], [False, True, False, False], [True, False, False, True]]
    torch.testing.assert_close(probs, torch.tensor([[0.6, 0, 0.4, 0], [0, 1.0, 0, 0], [0.7, 0, 0, 0.3]]))
    x = torch.arange(3, dtype=torch.float32)[:, None] * torch.ones(3, 4)
    permuted, pp, row_map = moe_permute_with_probs(x, probs, routing)
    assert permuted[:, 0].tolist() == [0.0, 2.0, 1.0, 0.0, 2.0]  # expert-major, token order kept inside an expert
    torch.testing.assert_close(pp, torch.tensor([0.6, 0.7, 1.0, 0.4, 0.3]))
    back = moe_unpermute_mask(permuted, row_map, merging_probs=probs, restore_shape=x.shape)
    torch.testing.assert_close(back, x * probs.sum(-1, keepdim=True))
