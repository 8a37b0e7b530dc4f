"""Fused expert block (one autograd node, accumulate-in-epilogue dgrads, caller-provided buffers) against the composed
grouped-linear / SiLU*mul path and an fp32 oracle."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(T=512, H=256, F=192, E=8, k=2, seed=0):
    from d9d_b200.kernel.moe import build_moe_layout, moe_permute

    torch.manual_seed(seed)
    x = torch.randn(T, H, device="cuda").bfloat16()
    ids = torch.stack([torch.randperm(E, device="cuda")[:k] for _ in range(T)])
    probs = torch.rand(T, k, device="cuda")
    layout = build_moe_layout(ids, E)
    xp, pp = moe_permute(x, probs, layout)
    w = [(torch.randn(E, a, b, device="cuda") * 0.05).bfloat16() for a, b in ((H, F), (H, F), (F, H))]
    return xp.detach(), pp.detach(), w, layout


def _oracle(xp, pp, w, layout):
    xs = xp.float().requires_grad_()
    ps = pp.float().requires_grad_()
    ws = [t.float().requires_grad_() for t in w]
    seg = layout.seg_offsets.tolist()
    out = torch.zeros(xp.shape[0], w[2].shape[2], device="cuda")
    rows = []
    for e in range(layout.num_experts):
        a, b = seg[e], seg[e + 1]
        if b > a:
            hcur = torch.nn.functional.silu(xs[a:b] @ ws[0][e]) * (xs[a:b] @ ws[1][e]) * ps[a:b, None]
            rows.append((a, b, hcur @ ws[2][e]))
    out = torch.cat([r[2] for r in rows]) if rows else out
    return xs, ps, ws, out, seg[-1]


def test_fused_expert_block_matches_oracle_and_composed_path():
    from d9d_b200.kernel.moe import grouped_linear, grouped_swiglu
    from d9d_b200.kernel.swiglu import silu_mul_probs

    xp, pp, w, layout = _setup()
    used = int(layout.seg_offsets[-1])
    dy = torch.randn(xp.shape[0], w[2].shape[2], device="cuda").bfloat16()

    def run(fused: bool):
        leaves = [xp.clone().requires_grad_(), pp.clone().requires_grad_()] + [t.clone().requires_grad_() for t in w]
        if fused:
            y = grouped_swiglu(leaves[0], leaves[1], *leaves[2:], layout)
        else:
            gate, up = grouped_linear(leaves[0], leaves[2], layout), grouped_linear(leaves[0], leaves[3], layout)
            y = grouped_linear(silu_mul_probs(gate, up, leaves[1]), leaves[4], layout)
        y[:used].backward(dy[:used])
        return y, [t.grad for t in leaves]

    y_f, g_f = run(True)
    y_c, g_c = run(False)
    torch.testing.assert_close(y_f[:used].float(), y_c[:used].float(), rtol=1e-2, atol=1e-2)
    for name, a, b in zip(("dx", "dprobs", "dwg", "dwu", "dwd"), g_f, g_c):
        a, b = a.float(), b.float()
        if name in ("dx", "dprobs"):
            a, b = a[:used], b[:used]
        scale = float(b.abs().max()) + 1e-6
        assert float((a - b).abs().max()) / scale < 3e-2, name

    xs, ps, ws, ref, _ = _oracle(xp, pp, w, layout)
    ref.backward(dy[:used].float())
    assert float((y_f[:used].float() - ref).abs().max()) / float(ref.abs().max()) < 3e-2
    scale = float(xs.grad.abs().max())
    assert float((g_f[0][:used].float() - xs.grad[:used]).abs().max()) / scale < 3e-2


def test_fused_expert_block_writes_into_caller_buffers_and_supports_split_backward():
    from d9d_b200.core.autograd import GLOBAL_GRAD_CONTEXT, GradDirection
    from d9d_b200.kernel.moe import grouped_swiglu

    xp, pp, w, layout = _setup(seed=3)
    used = int(layout.seg_offsets[-1])
    cap, H = xp.shape
    out_buf = torch.zeros(cap + 256, H, device="cuda", dtype=torch.bfloat16)
    dx_buf = torch.zeros(cap + 256, H, device="cuda", dtype=torch.bfloat16)
    dy = torch.randn(cap, H, device="cuda").bfloat16()

    leaves = [xp.clone().requires_grad_(), pp.clone().requires_grad_()] + [t.clone().requires_grad_() for t in w]
    y = grouped_swiglu(leaves[0], leaves[1], *leaves[2:], layout, out_buf, dx_buf)
    assert y.data_ptr() == out_buf.data_ptr()
    y.backward(dy)
    assert leaves[0].grad.data_ptr() == dx_buf.data_ptr() or torch.equal(leaves[0].grad[:used], dx_buf[:used])
    whole = [t.grad.clone() for t in leaves]

    leaves2 = [xp.clone().requires_grad_(), pp.clone().requires_grad_()] + [t.clone().requires_grad_() for t in w]
    y2 = grouped_swiglu(leaves2[0], leaves2[1], *leaves2[2:], layout)
    with GLOBAL_GRAD_CONTEXT.with_directions(GradDirection.inputs):
        dx, dp = torch.autograd.grad(y2, leaves2[:2], dy, retain_graph=True)
    with GLOBAL_GRAD_CONTEXT.with_directions(GradDirection.weight):
        dws = torch.autograd.grad(y2, leaves2[2:], dy, allow_unused=True)
    torch.testing.assert_close(dx[:used].float(), whole[0][:used].float(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(dp[:used], whole[1][:used], rtol=2e-2, atol=2e-2)
    for a, b in zip(dws, whole[2:]):
        torch.testing.assert_close(a.float(), b.float(), rtol=2e-2, atol=2e-2)
