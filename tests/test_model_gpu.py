"""Native-kernel model path (bf16, CUDA) against the pure-PyTorch fp32 oracle path of the same modules (CPU)."""

import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiny(moe: bool):
    from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode
    from d9d_b200.pipelining.api import PipelineStageInfo

    common = dict(hidden_size=256, intermediate_size=192, num_attention_heads=4, num_key_value_heads=2,
                  rms_norm_eps=1e-6, head_dim=64)
    vocab = dict(split_vocab_size={"regular": 1000, "special": 24}, split_vocab_order=["regular", "special"])
    if moe:
        from d9d_b200.module.model.qwen3_moe import (Qwen3MoEForCausalLM, Qwen3MoEForCausalLMParameters,
                                                     Qwen3MoELayerParameters, Qwen3MoEParameters)

        p = Qwen3MoEForCausalLMParameters(model=Qwen3MoEParameters(
            layer=Qwen3MoELayerParameters(num_experts=8, experts_top_k=2, **common), num_hidden_layers=2,
            rope_base=10000, max_position_ids=256, **vocab))
        return Qwen3MoEForCausalLM(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)
    from d9d_b200.module.model.qwen3_dense import (Qwen3DenseForCausalLM, Qwen3DenseForCausalLMParameters,
                                                   Qwen3DenseLayerParameters, Qwen3DenseParameters)

    p = Qwen3DenseForCausalLMParameters(model=Qwen3DenseParameters(
        layer=Qwen3DenseLayerParameters(**common), num_hidden_layers=2, rope_base=10000, max_position_ids=256, **vocab))
    return Qwen3DenseForCausalLM(p, PipelineStageInfo(0, 1), HiddenStatesAggregationMode.no, False)


def _cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return torch.nn.functional.cosine_similarity(a, b, dim=0).item()


@pytest.mark.parametrize("moe", [False, True])
def test_model_native_vs_oracle(moe):
    torch.manual_seed(0)
    ref = _tiny(moe)
    ref.reset_parameters()
    ref = ref.bfloat16().float()  # oracle holds the bf16-rounded weights in fp32
    ref.model.rope_provider.reset_parameters()
    gpu = copy.deepcopy(ref).bfloat16().cuda()
    gpu.model.rope_provider.reset_parameters()

    ids = torch.randint(0, 1024, (2, 128))
    labels = torch.randint(0, 1024, (2, 128))
    labels[0, :5] = -100
    pos = torch.arange(128)[None].expand(2, -1).contiguous()

    out_ref = ref(input_ids=ids, position_ids=pos, labels=labels)
    out_gpu = gpu(input_ids=ids.cuda(), position_ids=pos.cuda(), labels=labels.cuda())
    assert out_gpu["logps"].dtype == torch.float32
    l_ref, l_gpu = out_ref["logps"].sum() / 251, out_gpu["logps"].sum() / 251
    assert abs(l_ref.item() - l_gpu.item()) < 3e-2, (l_ref.item(), l_gpu.item())
    assert (out_gpu["logps"].cpu()[0, :5] == 0).all()
    l_ref.backward()
    l_gpu.backward()
    bad = []
    for (n, p_ref), (_, p_gpu) in zip(ref.named_parameters(), gpu.named_parameters()):
        assert p_gpu.grad is not None, n
        c = _cos(p_ref.grad, p_gpu.grad.cpu())
        ratio = (p_gpu.grad.float().norm().item() + 1e-12) / (p_ref.grad.norm().item() + 1e-12)
        if c < 0.98 or not (0.9 < ratio < 1.1):
            bad.append((n, c, ratio))
    assert not bad, bad


def test_smoke_entry():
    import __graft_entry__ as entry

    entry.smoke()


@pytest.mark.parametrize("grad_dtype", [torch.float32, torch.bfloat16])
def test_wgrad_accumulates_into_preallocated_grad(grad_dtype):
    """Opted-in parameters get ``grad += dW`` from the wgrad GEMM epilogue; result == autograd accumulation and the
    autograd still fires the post-accumulate hooks once per backward."""
    torch.manual_seed(1)
    base = _tiny(moe=True)
    base.reset_parameters()
    a = copy.deepcopy(base).bfloat16().cuda()
    b = copy.deepcopy(base).bfloat16().cuda()
    fired = {}
    for n, p in b.named_parameters():
        p.grad_dtype = grad_dtype
        p.grad = torch.zeros(p.shape, device="cuda", dtype=grad_dtype)
        p._d9d_fused_wgrad = True
        p.register_post_accumulate_grad_hook(lambda q, n=n: fired.__setitem__(n, fired.get(n, 0) + 1))
    for p in a.parameters():
        p.grad_dtype = grad_dtype
    ids = torch.randint(0, 1024, (2, 128), device="cuda")
    labels = torch.randint(0, 1024, (2, 128), device="cuda")
    pos = torch.arange(128, device="cuda")[None].expand(2, -1).contiguous()
    for _ in range(2):  # two accumulation rounds
        for m in (a, b):
            (m(input_ids=ids, position_ids=pos, labels=labels)["logps"].sum() / 256).backward()
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert fired.get(n) == 2, (n, fired.get(n))
        tol = 2e-2 if grad_dtype == torch.bfloat16 else 1e-4
        torch.testing.assert_close(pb.grad.float(), pa.grad.float(), rtol=tol, atol=tol * float(pa.grad.float().abs().max()),
                                   msg=lambda m: f"{n}: {m}")  # noqa: B023
