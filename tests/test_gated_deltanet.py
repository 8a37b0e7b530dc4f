import pytest
import torch

from d9d_b200.kernel.linear_attn import causal_conv1d_silu, chunk_gated_delta_rule, mamba_decay_gate, recurrent_gated_delta_rule
from d9d_b200.module.block.attention.linear import GatedDeltaNet, LogSigmoidDecayGateParameters, MambaDecayGateParameters


@pytest.mark.parametrize("seq", [1, 64, 150])
@pytest.mark.parametrize("l2norm", [True, False])
def test_chunked_delta_rule_matches_recurrence(seq, l2norm):
    torch.manual_seed(seq)
    B, H, dk, dv = 2, 3, 16, 24
    q, k = torch.randn(B, seq, H, dk), torch.randn(B, seq, H, dk) * (1.0 if l2norm else 0.3)
    v = torch.randn(B, seq, H, dv)
    g, beta = -torch.rand(B, seq, H) * 0.5, torch.rand(B, seq, H)
    want = recurrent_gated_delta_rule(q.double(), k.double(), v.double(), g.double(), beta.double(), use_qk_l2norm=l2norm)
    got = chunk_gated_delta_rule(q, k, v, g, beta, use_qk_l2norm=l2norm, chunk_size=64)
    torch.testing.assert_close(got.double(), want, rtol=1e-4, atol=1e-5)


def test_chunked_delta_rule_gradients():
    torch.manual_seed(1)
    B, S, H, dk, dv = 1, 70, 2, 8, 8
    args = [torch.randn(B, S, H, dk, dtype=torch.double), torch.randn(B, S, H, dk, dtype=torch.double),
            torch.randn(B, S, H, dv, dtype=torch.double), -torch.rand(B, S, H, dtype=torch.double) * 0.3,
            torch.rand(B, S, H, dtype=torch.double)]
    grads = []
    for fn in (recurrent_gated_delta_rule, lambda *a: chunk_gated_delta_rule(*a, chunk_size=32)):
        leaves = [a.clone().requires_grad_() for a in args]
        fn(*leaves).square().sum().backward()
        grads.append([leaf.grad for leaf in leaves])
    for a, b in zip(*grads):
        torch.testing.assert_close(a.float(), b.float(), rtol=2e-3, atol=2e-4)  # the chunked path computes in fp32


def test_causal_conv_is_causal_and_matches_definition():
    torch.manual_seed(0)
    x, w = torch.randn(2, 10, 6), torch.randn(6, 4)
    y = causal_conv1d_silu(x, w)
    ref = torch.zeros_like(x)
    for s in range(10):
        for j in range(4):
            src = s - 3 + j
            if src >= 0:
                ref[:, s] += w[:, j] * x[:, src]
    torch.testing.assert_close(y, torch.nn.functional.silu(ref), rtol=1e-5, atol=1e-5)
    x2 = x.clone()
    x2[:, 7:] += 1.0
    assert torch.equal(causal_conv1d_silu(x2, w)[:, :7], y[:, :7])


def test_mamba_gate_is_non_positive():
    g = mamba_decay_gate(torch.randn(4, 5, 3), torch.randn(3), torch.randn(3))
    assert (g <= 0).all() and g.dtype == torch.float32


@pytest.mark.parametrize("gate", [MambaDecayGateParameters(normalizer=16.0, dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4),
                                  LogSigmoidDecayGateParameters(normalizer=16.0)])
def test_gated_deltanet_module(gate):
    torch.manual_seed(0)
    m = GatedDeltaNet(hidden_size=32, num_query_key_heads=2, num_value_heads=4, head_qk_dim=8, head_v_dim=8, norm_eps=1e-6,
                      conv_size=4, decay_gate=gate)
    m.reset_parameters()
    x = torch.randn(2, 20, 32, requires_grad=True)
    mask = torch.ones(2, 20)
    mask[1, 15:] = 0
    y = m(x, mask)
    assert y.shape == x.shape
    y.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    # causality: changing the future must not change the past
    x2 = x.detach().clone()
    x2[:, 12:] += 0.5
    torch.testing.assert_close(m(x2, mask)[:, :12], y[:, :12].detach(), rtol=1e-4, atol=1e-5)
