"""MXFP8 recipe of ``Fp8Linear`` (forward, dgrad and wgrad quantised along their own reduction dimensions) on the GPU against
its CPU emulation.  Collected last on purpose: the composition was written after the round's GPU budget was spent."""

import pytest
import torch

from d9d_b200.kernel import fp8

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

@pytest.mark.xfail(strict=False, reason="composition of GPU-verified ops (quantize_mx, gemm_mxfp8) written after the round's GPU budget "
                                        "was spent: first executed by the round-end run")
def test_mxfp8_linear_function_matches_emulation():
    torch.manual_seed(0)
    x = torch.randn(4, 256, 512, device="cuda").bfloat16().requires_grad_()  # 1024 tokens
    w = (torch.randn(1024, 512, device="cuda") * 0.05).bfloat16().requires_grad_()
    y = fp8.fp8_linear(x, w, recipe="mx")
    g = torch.randn_like(y)
    y.backward(g)
    xc, wc = x.detach().cpu().requires_grad_(), w.detach().cpu().requires_grad_()
    yc = fp8.fp8_linear(xc, wc, recipe="mx")
    yc.backward(g.cpu())
    assert _rel(y.cpu(), yc) < 2e-2
    assert _rel(x.grad.cpu(), xc.grad) < 2e-2
    assert _rel(w.grad.cpu(), wc.grad) < 2e-2
    assert _rel(y, x.detach().float() @ w.detach().float().t()) < 0.06
