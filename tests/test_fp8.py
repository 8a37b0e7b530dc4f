"""FP8 recipes: the PyTorch emulation on CPU (scale rules, error bounds, autograd wiring)."""

import torch

from d9d_b200.kernel import fp8


def test_rowwise_quantisation_roundtrip():
    x = torch.randn(64, 128) * torch.logspace(-3, 2, 64)[:, None]
    q, s = fp8.quantize_rowwise_reference(x.bfloat16())
    back = q.float() * s[:, None]
    rel = (back - x.bfloat16().float()).abs().amax(1) / x.abs().amax(1)
    assert rel.max() < 0.07  # e4m3: 3 mantissa bits
    assert torch.isfinite(back).all() and (q.float().abs().amax(1) <= 448).all()


def test_mx_scale_exponent_is_tight_power_of_two():
    amax = torch.tensor([0.0, 448.0, 448.0001, 1.0, 3.4e-5, 7e4, 2.0**-20 * 448])
    e = fp8.mx_scale_exponent(amax)
    assert e[0] == -127
    for a, ee in zip(amax[1:], e[1:]):
        assert a / 2.0 ** ee.item() <= 448.0 * (1 + 1e-6)
        assert a / 2.0 ** (ee.item() - 1) > 448.0 * (1 - 1e-6)


def test_mx_roundtrip_and_block_layout():
    x = (torch.randn(200, 256) * torch.logspace(-2, 2, 8).repeat_interleave(32)[None]).bfloat16()
    q, e = fp8.quantize_mx_reference(x)
    back = fp8.dequantize_mx(q, e)
    assert ((back - x.float()).norm() / x.float().norm()) < 0.04
    # pack into the tensor-core block layout by hand and unpack again
    nb, kb = (200 + 255) // 256 * 2, 2
    sf = torch.zeros(nb, kb, 512, dtype=torch.uint8)
    for r in range(200):
        for g in range(8):
            sf[r // 128, g // 4, (r % 32) * 16 + ((r % 128) // 32) * 4 + g % 4] = e[r, g] + 127
    assert torch.equal(fp8.unpack_mx_scales(sf, 200, 256), e)


def test_fp8_linear_autograd_close_to_bf16():
    torch.manual_seed(0)
    x = torch.randn(4, 48, 64, requires_grad=True)
    w = (torch.randn(96, 64) * 0.1).requires_grad_()
    y = fp8.fp8_linear(x, w)
    y.float().square().mean().backward()
    x2, w2 = x.detach().clone().requires_grad_(), w.detach().clone().requires_grad_()
    y2 = x2 @ w2.t()
    y2.square().mean().backward()
    assert ((y.float() - y2).norm() / y2.norm()) < 0.06
    assert ((x.grad - x2.grad).norm() / x2.grad.norm()) < 0.08
    assert ((w.grad - w2.grad).norm() / w2.grad.norm()) < 0.08


def test_convert_linears():
    m = torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.ReLU(), torch.nn.Linear(64, 10))
    assert fp8.convert_linears_to_fp8(m) == 1 and isinstance(m[0], fp8.Fp8Linear) and type(m[2]) is torch.nn.Linear
    assert m(torch.randn(3, 32)).shape == (3, 10)


def test_mxfp8_linear_autograd_close_to_exact():
    torch.manual_seed(0)
    x = torch.randn(2, 64, 128, requires_grad=True)  # 128 tokens, in = 128
    w = (torch.randn(256, 128) * 0.1).requires_grad_()
    y = fp8.fp8_linear(x, w, recipe="mx")
    y.float().square().mean().backward()
    x2, w2 = x.detach().clone().requires_grad_(), w.detach().clone().requires_grad_()
    y2 = x2 @ w2.t()
    y2.square().mean().backward()
    assert ((y.float() - y2).norm() / y2.norm()) < 0.06
    assert ((x.grad - x2.grad).norm() / x2.grad.norm()) < 0.08
    assert ((w.grad - w2.grad).norm() / w2.grad.norm()) < 0.08
    import pytest

    with pytest.raises(ValueError):
        fp8.fp8_linear(torch.randn(3, 128), w, recipe="mx")  # 3 tokens: not a multiple of 128


def test_convert_linears_with_the_mx_recipe():
    m = torch.nn.Sequential(torch.nn.Linear(128, 256), torch.nn.ReLU(), torch.nn.Linear(256, 48))
    assert fp8.convert_linears_to_fp8(m, recipe="mx") == 1 and m[0].fp8_recipe == "mx" and type(m[2]) is torch.nn.Linear
    assert m(torch.randn(128, 128)).shape == (128, 48)
