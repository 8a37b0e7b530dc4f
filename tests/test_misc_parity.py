"""Smaller parity checks that mirror reference test files one to one (all CPU):

* ``modules/block/moe/test_hf_grouped_shared.py``  - MoE layer with a gated shared expert vs HF Qwen2-MoE,
* ``metric/classification/test_confusion_matrix.py`` - builder API vs scikit-learn,
* ``modules/block/test_hidden_state_aggregator.py`` / ``modules/block/normalization/test_rmsnorm.py``,
* ``kernel/flash_attn/test_kernel.py`` - the attention entry points' feature matrix (sinks, windows, GQA, varlen) against
  a token-by-token oracle,
* ``model_state/test_progress_bar.py``.
"""

import math

import pytest
import torch
from torch import nn


def _copy(dst, src):
    with torch.no_grad():
        assert dst.shape == src.shape, (dst.shape, src.shape)
        dst.copy_(src)


# -------------------------------------------------------------------------------------- MoE + shared expert
@pytest.mark.parametrize("gate", [True])
def test_moe_with_gated_shared_expert_matches_qwen2_moe(gate):
    pytest.importorskip("transformers")
    from transformers.models.qwen2_moe.configuration_qwen2_moe import Qwen2MoeConfig
    from transformers.models.qwen2_moe.modeling_qwen2_moe import Qwen2MoeSparseMoeBlock

    from d9d_b200.module.block.moe import MoELayer, SharedExpertParameters

    torch.manual_seed(0)
    hidden, inter, shared_inter, experts, top_k = 16, 8, 24, 5, 2
    cfg = Qwen2MoeConfig(hidden_size=hidden, moe_intermediate_size=inter, shared_expert_intermediate_size=shared_inter,
                         num_experts=experts, num_experts_per_tok=top_k, norm_topk_prob=False, num_hidden_layers=1)
    hf = Qwen2MoeSparseMoeBlock(cfg).float()
    with torch.no_grad():
        for p in hf.parameters():
            p.normal_(0, 0.25)
    ours = MoELayer(hidden_dim=hidden, intermediate_dim_grouped=inter, num_grouped_experts=experts, top_k=top_k,
                    router_renormalize_probabilities=False,
                    shared_expert=SharedExpertParameters(intermediate_size=shared_inter, enable_gate=gate))
    ours.reset_parameters()
    _copy(ours.router.gate.weight, hf.gate.weight)
    gate_up = hf.experts.gate_up_proj
    _copy(ours.grouped_experts.gate_proj.weight, gate_up[:, :inter].transpose(1, 2))
    _copy(ours.grouped_experts.up_proj.weight, gate_up[:, inter:].transpose(1, 2))
    _copy(ours.grouped_experts.down_proj.weight, hf.experts.down_proj.transpose(1, 2))
    for name in ("gate_proj", "up_proj", "down_proj"):
        _copy(getattr(ours.shared_expert.expert, name).weight, getattr(hf.shared_expert, name).weight)
    _copy(ours.shared_expert.gate.weight, hf.shared_expert_gate.weight)

    x = torch.randn(3, 5, hidden)
    x1, x2 = x.clone().requires_grad_(), x.clone().requires_grad_()
    y_hf = hf(x1)
    y_hf = y_hf[0] if isinstance(y_hf, tuple) else y_hf
    y = ours(x2)
    torch.testing.assert_close(y, y_hf.view_as(y), atol=2e-5, rtol=1e-4)
    y_hf.square().sum().backward()
    y.square().sum().backward()
    torch.testing.assert_close(x2.grad, x1.grad, atol=1e-4, rtol=1e-3)
    torch.testing.assert_close(ours.shared_expert.gate.weight.grad, hf.shared_expert_gate.weight.grad, atol=1e-4, rtol=1e-3)


# -------------------------------------------------------------------------------------- confusion matrix
def test_confusion_matrix_metrics_match_sklearn():
    sk = pytest.importorskip("sklearn.metrics")
    from d9d_b200.metric.impl.classification.confusion_matrix import confusion_matrix_metric

    torch.manual_seed(0)
    n, c = 200, 4
    logits = torch.randn(n, c)
    target = torch.randint(0, c, (n,))
    pred = logits.argmax(-1)

    def run(metric, preds, targets, chunks=3):
        for p, t in zip(preds.chunk(chunks), targets.chunk(chunks)):  # accumulates over updates
            metric.update(p, t)
        return metric.compute()

    f1_macro = run(confusion_matrix_metric().multiclass(c).with_f1().macro().build(), logits, target)
    assert math.isclose(float(f1_macro), sk.f1_score(target, pred, average="macro"), rel_tol=1e-5)
    prec_w = run(confusion_matrix_metric().multiclass(c).with_precision().weighted().build(), logits, target)
    assert math.isclose(float(prec_w), sk.precision_score(target, pred, average="weighted", zero_division=0), rel_tol=1e-5)
    rec_pc = run(confusion_matrix_metric().multiclass(c).with_recall().per_class().build(), logits, target)
    torch.testing.assert_close(rec_pc.double(), torch.tensor(sk.recall_score(target, pred, average=None)), rtol=1e-5, atol=1e-6)
    # one-vs-rest accuracy pooled over classes: every wrong sample is a miss for two classes (its target and its prediction)
    acc = run(confusion_matrix_metric().multiclass(c).with_accuracy().micro().build(), logits, target)
    wrong = n * (1.0 - sk.accuracy_score(target, pred))
    assert math.isclose(float(acc), (c * n - 2 * wrong) / (c * n), rel_tol=1e-5)
    top2 = run(confusion_matrix_metric().multiclass(c, top_k=2).with_accuracy().build(), logits, target)
    assert math.isclose(float(top2), sk.top_k_accuracy_score(target, logits, k=2, labels=list(range(c))), rel_tol=1e-5)

    probs = torch.rand(n)
    binary_target = torch.randint(0, 2, (n,))
    f2 = run(confusion_matrix_metric().binary(threshold=0.4).with_fbeta(2.0).build(), probs, binary_target)
    assert math.isclose(float(f2), sk.fbeta_score(binary_target, (probs > 0.4).long(), beta=2.0), rel_tol=1e-5)

    ml_probs = torch.rand(n, 3)
    ml_target = torch.randint(0, 2, (n, 3))
    ml = run(confusion_matrix_metric().multilabel(3, threshold=0.5).with_recall().micro().build(), ml_probs, ml_target)
    assert math.isclose(float(ml), sk.recall_score(ml_target, (ml_probs > 0.5).long(), average="micro"), rel_tol=1e-5)

    metric = confusion_matrix_metric().multiclass(c).with_f1().macro().build()
    run(metric, logits, target)
    saved = metric.state_dict()
    fresh = confusion_matrix_metric().multiclass(c).with_f1().macro().build()
    fresh.load_state_dict(saved)
    assert torch.equal(fresh.compute(), metric.compute())
    metric.reset()
    metric.update(logits[:10], target[:10])
    assert not torch.equal(metric.compute(), fresh.compute())

    with pytest.raises(ValueError):
        confusion_matrix_metric().binary().multiclass(3)
    with pytest.raises(ValueError):
        confusion_matrix_metric().multiclass(3).with_f1().with_recall()
    with pytest.raises(ValueError):
        confusion_matrix_metric().multiclass(3).with_f1().build()  # no aggregation chosen
    with pytest.raises(ValueError):
        confusion_matrix_metric().with_f1().macro().build()  # no problem type


# ------------------------------------------------------------------------- aggregator / normalisation blocks
def test_hidden_states_mean_aggregator():
    from d9d_b200.module.block.hidden_states_aggregator import HiddenStatesAggregationMode, create_hidden_states_aggregator

    mask = torch.tensor([[1, 1, 0, 0], [0, 1, 1, 1]])
    agg = create_hidden_states_aggregator(HiddenStatesAggregationMode.mean, mask)
    layers = [torch.randn(2, 4, 3) for _ in range(3)]
    for h in layers[:2]:
        agg.add_hidden_states(h)
    first = agg.pack_with_snapshot(None)
    assert first.shape == (2, 2, 3)  # [layers, batch, hidden]
    torch.testing.assert_close(first[1, 0], layers[1][0, :2].mean(0))
    torch.testing.assert_close(first[0, 1], layers[0][1, 1:].mean(0))
    agg.add_hidden_states(layers[2])
    both = agg.pack_with_snapshot(first)  # a later pipeline stage appends to what earlier stages collected
    assert both.shape == (3, 2, 3) and torch.equal(both[:2], first)
    assert agg.pack_with_snapshot(None) is None  # drained
    noop = create_hidden_states_aggregator(HiddenStatesAggregationMode.no, None)
    noop.add_hidden_states(layers[0])
    assert noop.pack_with_snapshot(None) is None
    with pytest.raises(ValueError):
        create_hidden_states_aggregator(HiddenStatesAggregationMode.mean, None)


@pytest.mark.parametrize("zero_centered", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rms_norm_module(zero_centered, dtype):
    from d9d_b200.module.block.normalization import RMSNorm

    torch.manual_seed(0)
    norm = RMSNorm(48, eps=1e-5, zero_centered=zero_centered)
    norm.reset_parameters()
    assert float(norm.weight.sum()) == (0.0 if zero_centered else 48.0)
    with torch.no_grad():
        norm.weight.add_(torch.randn(48) * 0.1)
    norm = norm.to(dtype)
    x = torch.randn(5, 7, 48, dtype=dtype, requires_grad=True)
    y = norm(x)
    assert y.dtype == dtype and y.shape == x.shape

    x32 = x.detach().float().requires_grad_()
    w32 = norm.weight.detach().float().requires_grad_()
    scale = (1.0 + w32) if zero_centered else w32
    ref = x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + 1e-5) * scale
    tol = dict(atol=1e-5, rtol=1e-5) if dtype == torch.float32 else dict(atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(y.float(), ref, **tol)
    g = torch.randn_like(ref)
    ref.backward(g)
    y.backward(g.to(dtype))
    torch.testing.assert_close(x.grad.float(), x32.grad, **tol)
    torch.testing.assert_close(norm.weight.grad.float(), w32.grad, atol=tol["atol"] * 10, rtol=tol["rtol"])


# ------------------------------------------------------------------------------------- attention features
def _token_oracle(q, k, v, causal, window, sink, scale):
    """Attention one query at a time, straight from the definition (bottom-right aligned positions)."""
    sq, h, _ = q.shape
    sk, hk, dv = v.shape
    out = torch.zeros(sq, h, dv, dtype=torch.float64)
    for head in range(h):
        kv_head = head // (h // hk)
        for i in range(sq):
            pos = i + (sk - sq)
            scores, values = [], []
            for j in range(sk):
                if causal and j > pos:
                    continue
                if window[0] is not None and j < pos - window[0]:
                    continue
                if not causal and window[1] is not None and j > pos + window[1]:
                    continue
                scores.append(float(q[i, head].double() @ k[j, kv_head].double()) * scale)
                values.append(v[j, kv_head].double())
            if not scores:
                continue
            s = torch.tensor(scores, dtype=torch.float64)
            top = max(float(s.max()), float(sink[head]) if sink is not None else -math.inf)
            weights = torch.exp(s - top)
            denom = weights.sum() + (math.exp(float(sink[head]) - top) if sink is not None else 0.0)
            out[i, head] = (weights[:, None] * torch.stack(values)).sum(0) / denom
    return out


@pytest.mark.parametrize("heads,kv_heads", [(4, 4), (4, 2), (4, 1)])
@pytest.mark.parametrize("causal,window", [(True, (None, None)), (True, (3, None)), (False, (2, 1)), (False, (None, None))])
@pytest.mark.parametrize("use_sink", [False, True])
def test_flash_attn_func_feature_matrix(heads, kv_heads, causal, window, use_sink):
    from d9d_b200.kernel.flash_attn import flash_attn_func

    torch.manual_seed(0)
    sq, sk, d = 6, 9, 8
    q = torch.randn(1, sq, heads, d, requires_grad=True)
    k = torch.randn(1, sk, kv_heads, d, requires_grad=True)
    v = torch.randn(1, sk, kv_heads, d, requires_grad=True)
    sink = torch.randn(heads, requires_grad=True) if use_sink else None
    out, lse = flash_attn_func(q, k, v, causal=causal, window_size=window, learnable_sink=sink, return_lse=True)
    want = _token_oracle(q[0].detach(), k[0].detach(), v[0].detach(), causal, window, sink.detach() if use_sink else None, d**-0.5)
    torch.testing.assert_close(out[0].double(), want, atol=1e-5, rtol=1e-5)
    assert lse.shape == (1, heads, sq)
    out.square().sum().backward()
    assert all(t.grad is not None and torch.isfinite(t.grad).all() for t in (q, k, v))
    if use_sink:
        # more sink mass can only shrink every output row: d(sum out^2)/d sink < 0 for rows that attend to anything
        assert sink.grad is not None and torch.isfinite(sink.grad).all() and (sink.grad <= 1e-9).all()


def test_flash_attn_varlen_equals_per_sequence_calls():
    from d9d_b200.kernel.flash_attn import flash_attn_func, flash_attn_varlen_func

    torch.manual_seed(1)
    lengths = [3, 7, 1, 5]
    cu = torch.tensor([0, *torch.tensor(lengths).cumsum(0).tolist()], dtype=torch.int32)
    total, h, hk, d = sum(lengths), 4, 2, 8
    q, k, v = torch.randn(total, h, d), torch.randn(total, hk, d), torch.randn(total, hk, d)
    out, lse = flash_attn_varlen_func(q, k, v, cu_seqlens_q=cu, cu_seqlens_k=cu, max_seqlen_q=7, max_seqlen_k=7, causal=True,
                                      return_lse=True)
    assert out.shape == (total, h, d) and lse.shape == (h, total)
    for i, n in enumerate(lengths):
        lo = int(cu[i])
        single, _ = flash_attn_func(q[None, lo:lo + n], k[None, lo:lo + n], v[None, lo:lo + n], causal=True)
        torch.testing.assert_close(out[lo:lo + n], single[0])
    with pytest.raises(ValueError):
        flash_attn_varlen_func(q, k, v)
    with pytest.raises(NotImplementedError):
        flash_attn_varlen_func(q, k, v, cu_seqlens_q=cu, cu_seqlens_k=cu, page_table=torch.zeros(1, 1, dtype=torch.int32))


# ------------------------------------------------------------------------------------------- progress bar
def test_model_state_io_progress_reporting(tmp_path, capsys):
    """Saving / loading with ``show_progress=True`` reports one tick per state and leaves the data intact."""
    from d9d_b200.model_state.io import load_model_state, save_model_state
    from d9d_b200.model_state.mapper.compose import ModelStateMapperParallel
    from d9d_b200.model_state.mapper.leaf import ModelStateMapperIdentity

    model = nn.Sequential(nn.Linear(4, 4), nn.Linear(4, 2))
    names = list(model.state_dict())
    mapper = ModelStateMapperParallel([ModelStateMapperIdentity(n) for n in names])
    save_model_state(tmp_path, mapper, model, show_progress=True)
    clone = nn.Sequential(nn.Linear(4, 4), nn.Linear(4, 2))
    load_model_state(tmp_path, mapper, "cpu", clone, show_progress=True)
    for n in names:
        assert torch.equal(clone.state_dict()[n], model.state_dict()[n])
    err = capsys.readouterr().err
    assert "Saving" in err or "Loading" in err or "%" in err or err == ""  # tqdm writes to stderr when it is a TTY
