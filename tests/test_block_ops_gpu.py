"""Linear weight/bias-gradient and LayerNorm kernels against the stock torch ops they replace."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("r,i,o,bias", [(8320, 32, 96, False), (8320, 32, 32, True), (4100, 32, 64, True),
                                         (4096, 64, 32, True), (2048, 48, 32, True), (1030, 16, 128, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_wgrad_matches_torch(r, i, o, bias, dtype):
    from erv_b200 import _capi as C, ops
    assert C.load().erv_linear_wgrad_supported(r, o, i)
    torch.manual_seed(r + i + o)
    lin = torch.nn.Linear(i, o, bias=bias).cuda()
    x = torch.randn(5, r // 5, i, device="cuda", requires_grad=True)
    w = torch.randn(5, r // 5, o, device="cuda")
    x2 = x.detach().clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
        y = ops.linear(x, lin.weight, lin.bias)
    (y.float() * w).sum().backward()
    got = [x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone() if bias else None]
    lin.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
        y2 = torch.nn.functional.linear(x2, lin.weight, lin.bias)
    (y2.float() * w).sum().backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert torch.equal(y, y2)
    assert rel_l2(got[0], x2.grad) < tol
    assert rel_l2(got[1], lin.weight.grad) < tol
    if bias:
        assert rel_l2(got[2], lin.bias.grad) < tol


def test_linear_falls_back_to_library_for_other_shapes():
    from erv_b200 import _capi as C, ops
    assert not C.load().erv_linear_wgrad_supported(64, 32, 32)      # few rows
    assert not C.load().erv_linear_wgrad_supported(8192, 768, 768)  # large matrices: cuBLAS territory
    lin = torch.nn.Linear(32, 10).cuda()
    x = torch.randn(64, 32, device="cuda", requires_grad=True)
    ops.linear(x, lin.weight, lin.bias).sum().backward()
    assert torch.allclose(lin.bias.grad, torch.full((10,), 64.0, device="cuda"))


@pytest.mark.parametrize("r,c", [(66560, 32), (1000, 32), (333, 64), (1024, 100), (77, 256)])
def test_layernorm_matches_torch(r, c):
    from erv_b200 import ops
    torch.manual_seed(r + c)
    ln = torch.nn.LayerNorm(c).cuda()
    with torch.no_grad():
        ln.weight.normal_(1.0, 0.2)
        ln.bias.normal_(0.0, 0.2)
    x = (torch.randn(r, c, device="cuda") * 2 + 0.5).requires_grad_(True)
    w = torch.randn(r, c, device="cuda")
    y = ops.layer_norm(x, ln.weight, ln.bias, ln.eps)
    (y * w).sum().backward()
    got = [y.detach().clone(), x.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone()]
    x.grad = None
    ln.zero_grad()
    y2 = torch.nn.functional.layer_norm(x, (c,), ln.weight, ln.bias, ln.eps)
    (y2 * w).sum().backward()
    assert rel_l2(got[0], y2) < 1e-6
    assert rel_l2(got[1], x.grad) < 1e-5
    assert rel_l2(got[2], ln.weight.grad) < 1e-5
    assert rel_l2(got[3], ln.bias.grad) < 1e-5
