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


# ---- fused block kernels (csrc/erv_block_fused.cu) ---------------------------------------------------------------------
def _make_block(dropout=0.0):
    from erv_b200 import FAVORPlusAttention
    from erv_b200.vit import UnifiedTransformerBlock
    torch.manual_seed(11)
    blk = UnifiedTransformerBlock(32, FAVORPlusAttention(32, 2, dropout=dropout, num_features=64), None, 64, dropout)
    with torch.no_grad():  # non-trivial LayerNorm affine parameters and biases
        for p in blk.parameters():
            if p.dim() == 1:
                p.normal_(0.5, 0.3)
    return blk.to("cuda")


@pytest.mark.parametrize("shape", [(3, 65, 32), (1, 5, 32), (16, 197, 32)])
def test_fused_block_matches_unfused(shape, monkeypatch):
    """The two fused kernels (+ their backward) against the same block run op by op (library GEMMs, ATen GELU, ...)."""
    from erv_b200 import ops
    blk = _make_block().eval()
    x = torch.randn(*shape, device="cuda")
    w = torch.randn(*shape, device="cuda")
    res = []
    for fused in (True, False):
        monkeypatch.setattr(ops, "FUSED_BLOCK", fused)
        blk.zero_grad()
        xi = x.clone().requires_grad_(True)
        y = blk(xi)
        (y * w).sum().backward()
        res.append((y.detach(), xi.grad, {k: p.grad.clone() for k, p in blk.named_parameters()}))
    (y1, dx1, g1), (y0, dx0, g0) = res
    assert rel_l2(y1, y0) < 1e-5 and rel_l2(dx1, dx0) < 2e-5
    for k in g0:
        assert rel_l2(g1[k], g0[k]) < 5e-5, k


def test_fused_block_dropout_is_consistent():
    """Training-mode dropout: keep rate, forward/backward use the same masks (directional derivative), and a new
    seed gives new masks."""
    from erv_b200 import ops
    torch.manual_seed(3)
    a = torch.randn(512, 65, 32, device="cuda")
    x = torch.zeros(512, 65, 32, device="cuda")
    eye, zero = torch.eye(32, device="cuda"), torch.zeros(32, device="cuda")
    ones = torch.ones(32, device="cuda")
    w1, b1 = torch.zeros(64, 32, device="cuda"), torch.zeros(64, device="cuda")
    w2 = torch.zeros(32, 64, device="cuda")
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    # identity projection, zero MLP: y = drop(a)
    y = ops.block_mlp(a, x, eye, zero, ones, zero, w1, b1, w2, zero, 1e-5, 0.25, seed)
    kept = (y != 0).float().mean().item()
    assert abs(kept - 0.75) < 0.01
    assert torch.allclose(y[y != 0], (a / 0.75)[y != 0], rtol=1e-4, atol=1e-6)  # split-bf16 tensor-core product
    y2 = ops.block_mlp(a, x, eye, zero, ones, zero, w1, b1, w2, zero, 1e-5, 0.25, seed + 1)
    assert ((y != 0) != (y2 != 0)).float().mean().item() > 0.2
    # full block: gradient along a random direction equals the central difference with the masks held fixed
    blk = _make_block(dropout=0.2).train()
    att = blk.attention
    args = lambda t: (t, xin, att.proj.weight, att.proj.bias, blk.norm2.weight, blk.norm2.bias, blk.mlp[0].weight,  # noqa: E731
                      blk.mlp[0].bias, blk.mlp[3].weight, blk.mlp[3].bias, 1e-5, 0.2, seed)
    xin = torch.randn(64, 65, 32, device="cuda")
    a0 = torch.randn(64, 65, 32, device="cuda", dtype=torch.float64).float().requires_grad_(True)
    v = torch.randn_like(a0)
    wgt = torch.randn_like(a0)
    (ops.block_mlp(*args(a0)) * wgt).sum().backward()
    eps = 1e-2
    with torch.no_grad():
        fd = ((ops.block_mlp(*args(a0 + eps * v)) - ops.block_mlp(*args(a0 - eps * v))) * wgt).double().sum() / (2 * eps)
    an = (a0.grad * v).double().sum()
    assert abs(float(fd - an)) / abs(float(an)) < 2e-2


def test_fused_gradient_accumulation_matches_autograd(monkeypatch):
    """ops.GRAD_INPLACE (used by the Trainer): the backward kernels add into existing .grad buffers; two backward
    passes must accumulate exactly like autograd's AccumulateGrad does."""
    from erv_b200 import ops
    blk = _make_block().eval()
    xs = [torch.randn(4, 65, 32, device="cuda") for _ in range(2)]
    got = []
    for inplace in (False, True):
        monkeypatch.setattr(ops, "GRAD_INPLACE", inplace)
        for p in blk.parameters():
            p.grad = torch.zeros_like(p)
        for x in xs:
            blk(x).square().sum().backward()
        got.append({k: p.grad.clone() for k, p in blk.named_parameters()})
    for k in got[0]:
        assert rel_l2(got[1][k], got[0][k]) < 1e-6, k


def test_trainer_with_flat_parameters_and_kerple():
    """KERPLE's rel_pos_bias has 2(2N-1) elements: the flat parameter buffer must keep every parameter 16-byte aligned
    for the fused block kernels (regression: misaligned address at BASELINE configs 3 and 5)."""
    from erv_b200 import MNIST_CONFIG, create_model
    from erv_b200.train import Trainer
    torch.manual_seed(0)
    model = create_model("performer_relu_most_general", MNIST_CONFIG).to("cuda").train()
    tr = Trainer(model, lr=1e-3, use_graph=False)
    for p in tr.params:
        assert p.data_ptr() % 16 == 0 and p.grad.data_ptr() % 16 == 0
    img, lab = torch.randn(16, 1, 28, 28, device="cuda"), torch.randint(0, 10, (16,), device="cuda")
    losses = [float(tr.step(img, lab)) for _ in range(5)]
    assert all(l == l and l < 100 for l in losses) and losses[-1] < losses[0]


def test_prefetched_steps_equal_direct_steps():
    """Trainer.prefetch / step_prefetched (host batches copied on a side stream) train exactly like Trainer.step."""
    from erv_b200 import MNIST_CONFIG, create_model
    from erv_b200.train import Trainer
    batches = [(torch.randn(8, 1, 28, 28).pin_memory(), torch.randint(0, 10, (8,)).pin_memory()) for _ in range(4)]
    losses = []
    for mode in ("direct", "prefetched"):
        torch.manual_seed(0)
        tr = Trainer(create_model("performer_favor", MNIST_CONFIG, dropout=0.0).to("cuda").train(), use_graph=True)
        out = []
        if mode == "direct":
            for b in batches:
                out.append(float(tr.step(*b)))
        else:
            tr.prefetch(*batches[0])
            for i in range(len(batches)):
                l = tr.step_prefetched()
                if i + 1 < len(batches):
                    tr.prefetch(*batches[i + 1])
                out.append(float(l))
        losses.append(out)
    assert losses[0] == losses[1]


# ---- the two ends of the ViT (csrc/erv_embed_head.cu) --------------------------------------------------------------------
@pytest.mark.parametrize("cfg,patch,bsz", [("mnist", 7, 5), ("cifar", 4, 6), ("cifar", 8, 3)])
def test_fused_embed_and_head_loss_match_library_path(cfg, patch, bsz, monkeypatch):
    """model.loss(images, labels) with the fused embedding / head+loss kernels against the same model run op by op:
    loss and every parameter gradient (MNIST patch dim 49 is not a multiple of 8; CIFAR patch 8 is the 192-wide case)."""
    from erv_b200 import CIFAR10_CONFIG, MNIST_CONFIG, create_model, ops
    torch.manual_seed(4)
    base = MNIST_CONFIG if cfg == "mnist" else CIFAR10_CONFIG
    model = create_model("performer_favor", base, patch_size=patch, dropout=0.0).to("cuda").eval()
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.normal_(0.3, 0.3)
    s, c = (28, 1) if cfg == "mnist" else (32, 3)
    img, lab = torch.randn(bsz, c, s, s, device="cuda"), torch.randint(0, 10, (bsz,), device="cuda")
    res = []
    for fused in (True, False):
        monkeypatch.setattr(ops, "FUSED_BLOCK", fused)
        monkeypatch.setattr(ops, "FUSED_EMBED", "always" if fused else False)
        model.zero_grad()
        loss = model.loss(img, lab)
        loss.backward()
        res.append((loss.detach(), {k: p.grad.clone() for k, p in model.named_parameters()}))
    assert abs(float(res[0][0]) - float(res[1][0])) < 1e-5
    for k in res[1][1]:
        assert rel_l2(res[0][1][k], res[1][1][k]) < 5e-5, k
    monkeypatch.setattr(ops, "FUSED_BLOCK", True)
    logits = model(img)  # the plain forward keeps its meaning
    assert abs(float(torch.nn.functional.cross_entropy(logits, lab).detach()) - float(res[0][0])) < 1e-5


@pytest.mark.parametrize("name", ["performer_favor_circulant", "baseline_rope", "performer_relu_most_general"])
def test_fused_path_under_bf16_autocast(name, monkeypatch):
    """Under bf16 autocast the fused block / embedding / head kernels stay in fp32 and hand the attention core bf16 qkv:
    loss and gradients agree with the op-by-op autocast path within the bf16 budget (2e-2), and are closer to fp32."""
    from erv_b200 import CIFAR10_CONFIG, create_model, ops
    torch.manual_seed(6)
    model = create_model(name, CIFAR10_CONFIG, patch_size=4, dropout=0.0).to("cuda").eval()
    img, lab = torch.randn(8, 3, 32, 32, device="cuda"), torch.randint(0, 10, (8,), device="cuda")
    res = {}
    for key, fused, ac in (("fused_bf16", True, True), ("plain_bf16", False, True), ("fp32", True, False)):
        monkeypatch.setattr(ops, "FUSED_BLOCK", fused)
        model.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            loss = model.loss(img, lab)
        loss.backward()
        res[key] = (float(loss), {k: p.grad.float().clone() for k, p in model.named_parameters()})
    assert abs(res["fused_bf16"][0] - res["fp32"][0]) < 2e-2 and abs(res["plain_bf16"][0] - res["fp32"][0]) < 5e-2
    worst_f = max(rel_l2(res["fused_bf16"][1][k], res["fp32"][1][k]) for k in res["fp32"][1])
    worst_p = max(rel_l2(res["plain_bf16"][1][k], res["fp32"][1][k]) for k in res["fp32"][1])
    assert worst_f < 4e-2 and worst_f <= worst_p * 1.5, (worst_f, worst_p)
