"""GPU parity: the CUDA path (through the plugin classes -> ctypes -> C ABI) against
  (1) tests/golden/*.npz produced by the unmodified reference, and
  (2) oracle/erv_oracle.py on seeded inputs at sizes the oracle finishes in seconds.
Tolerances are BASELINE.json's: 1e-4 relative (fp32), 2e-2 (bf16); "relative" = ||a-b|| / ||b||, and next to it
max|a-b| / max|b| within conftest.MAX_FACTOR x the same tolerance (assert_close checks both)."""
import pytest
import torch

from conftest import ATTN_KIND, RPE_KIND, assert_close, golden_files, load_golden, parse_attn_case, rel_l2

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-4
TOL_BF16 = 2e-2
DEV = "cuda"


def _build(g, a, r):
    from erv_b200 import ATTENTION_REGISTRY, RPE_REGISTRY
    heads = int(g["heads"])
    _, n, dim = g["x"].shape
    kw = {"num_features": g["attn.omega"].shape[-1]} if a != "softmax" else {}
    attn = ATTENTION_REGISTRY[a](dim=dim, heads=heads, dropout=0.0, **kw)
    attn.load_state_dict({k[5:]: v for k, v in g.items() if k.startswith("attn.")})
    rpe = None
    if r is not None:
        rpe = RPE_REGISTRY[r](num_patches=n, dim=dim, heads=heads)
        rpe.load_state_dict({k[4:]: v for k, v in g.items() if k.startswith("rpe.")})
        rpe = rpe.to(DEV)
    return attn.to(DEV).eval(), rpe


@pytest.mark.parametrize("fname", golden_files("attn_"))
def test_attention_fp32_matches_reference(fname):
    _, a, r = parse_attn_case(fname)
    g = load_golden(fname)
    attn, rpe = _build(g, a, r)
    x = g["x"].to(DEV).requires_grad_(True)
    out = attn(x, rpe=rpe)
    assert out.shape == x.shape and out.dtype == torch.float32
    assert_close(out, g["out"], TOL_F32, "out")
    (out * g["cotangent"].to(DEV)).sum().backward()
    assert_close(x.grad, g["dx"], TOL_F32, "dx")
    for k, p in attn.named_parameters():
        assert_close(p.grad, g["grad.attn." + k], TOL_F32, k)
    if rpe is not None:
        for k, p in rpe.named_parameters():
            assert_close(p.grad, g["grad.rpe." + k], TOL_F32, k)


@pytest.mark.parametrize("fname", golden_files("attn_"))
def test_attention_bf16_autocast_within_tolerance(fname):
    _, a, r = parse_attn_case(fname)
    g = load_golden(fname)
    attn, rpe = _build(g, a, r)
    x = g["x"].to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = attn(x, rpe=rpe)
    assert out.dtype == torch.bfloat16
    (out.float() * g["cotangent"].to(DEV)).sum().backward()
    # against the reference's own autocast run, and against its fp32 run (bf16 noise floor: SURVEY.md 8(c))
    assert_close(out, g["out_bf16"], TOL_BF16, "out vs autocast reference")
    assert_close(out, g["out"], TOL_BF16, "out vs fp32 reference")
    assert_close(x.grad, g["dx_bf16"], 2 * TOL_BF16, "dx vs autocast reference")
    assert_close(x.grad, g["dx"], 2 * TOL_BF16, "dx vs fp32 reference")


@pytest.mark.parametrize("fname", golden_files("model_"))
def test_model_matches_reference(fname):
    from erv_b200 import CIFAR10_CONFIG, MNIST_CONFIG, create_model
    g = load_golden(fname)
    name = fname[len("model_"):-4]
    if name.startswith("cifar_"):
        model = create_model("performer_favor", CIFAR10_CONFIG, attention_config={"num_features": 256},
                             patch_size=4, dropout=0.0)
    else:
        model = create_model(name, MNIST_CONFIG, dropout=0.0)
    model.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("sd.")})
    model = model.to(DEV).eval()
    logits = model(g["images"].to(DEV))
    assert_close(logits, g["logits"], TOL_F32, "logits")
    loss = torch.nn.functional.cross_entropy(logits, g["labels"].to(DEV))
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    loss.backward()
    params = dict(model.named_parameters())
    for k, v in g.items():
        if k.startswith("grad."):
            assert_close(params[k[5:]].grad, v, 2 * TOL_F32, k)


def test_units_match_reference():
    from erv_b200 import CirculantStringRPE, FAVORPlusAttention, ReLUAttention, RoPE, ops
    from erv_b200.rpe.fft_utils import fft_toeplitz_matmul
    g = {k: v.to(DEV) for k, v in load_golden("units.npz").items()}
    assert rel_l2(fft_toeplitz_matmul(g["toep.c"], g["toep.x"]), g["toep.y"]) < TOL_F32
    assert rel_l2(fft_toeplitz_matmul(g["toepb.c"], g["toepb.x"]), g["toepb.y"]) < TOL_F32
    fav, rel = FAVORPlusAttention(32, 2).to(DEV), ReLUAttention(32, 2).to(DEV)
    phi = fav._compute_phi_positive(g["feat.x"], g["feat.omega_favor"])
    assert (phi >= 0).all() and rel_l2(phi, g["feat.phi_favor"]) < TOL_F32
    assert rel_l2(rel._compute_relu_features(g["feat.x"], g["feat.omega_relu"]), g["feat.phi_relu"]) < TOL_F32
    rope = RoPE(10, 32, 2).to(DEV)
    assert torch.equal(rope.cos_cached, g["rope.cos"]) and torch.equal(rope.sin_cached, g["rope.sin"])
    q, k = rope.apply_rotary_emb(g["feat.x"], g["rope.k"])
    assert rel_l2(q, g["rope.q_out"]) < 1e-6 and rel_l2(k, g["rope.k_out"]) < 1e-6
    cos, sin = ops.rope_table(10000.0, 10, 16, DEV)  # device-side builder agrees with the host caches
    assert (cos - g["rope.cos"]).abs().max() < 1e-6 and (sin - g["rope.sin"]).abs().max() < 1e-6
    circ = CirculantStringRPE(10, 32, 2).to(DEV)
    with torch.no_grad():
        circ.circulant_coeffs.copy_(g["circ.coeffs"])
    qc, kc = circ.apply_circulant_string(g["circ.q"], g["circ.k"])
    assert rel_l2(qc, g["circ.q_out"]) < 1e-5 and rel_l2(kc, g["circ.k_out"]) < 1e-5
    assert torch.equal(qc[:, :, 0], g["circ.q"][:, :, 0])  # CLS row untouched
    pr = circ.apply_rotation(g["circ.q"][:, :, 1:], circ.patch_positions)
    assert rel_l2(pr, g["circ.q_out"][:, :, 1:]) < 1e-5


# ---- oracle at BASELINE shapes (seeded inputs, sizes the CPU oracle finishes in seconds) -------------------
ORACLE_CASES = [
    # attention, rpe, B, N, dim, heads, M
    ("favor_plus", None, 16, 65, 32, 2, 256),             # BASELINE config 2
    ("relu", "most_general", 16, 65, 32, 2, 44),          # config 3' (p4)
    ("relu", "most_general", 32, 17, 32, 2, 44),          # config 3 (p8)
    ("favor_plus", "circulant_string", 4, 197, 32, 2, 44),  # config 4
    ("softmax", "rope", 4, 197, 32, 2, None),             # config 4
    ("softmax", None, 128, 17, 32, 2, None),              # config 1
    ("favor_plus", "most_general", 1, 257, 32, 2, 44),    # config 5, shortened (N=257, several key tiles)
    ("favor_plus", "most_general", 1, 4097, 32, 2, 44),   # config 5 at its real shape (dense oracle route: 2 x 16.8 M scores)
    ("favor_plus", "rope", 2, 197, 768, 12, 64),          # ViT-B fixture dims, Dh=64
    ("relu", "circulant_string", 3, 50, 64, 8, 24),       # Dh=8
    # short sequences: two (batch, head) pairs per tensor-core tile (erv_linattn_tc2.cu)
    ("favor_plus", "rope", 3, 65, 32, 2, 44),             # odd batch, lone token, 64-feature instance
    ("relu", "circulant_string", 5, 65, 32, 2, 100),      # 128-feature instance
    ("favor_plus", None, 4, 50, 32, 2, 256),              # no lone token, ragged tile halves
    ("relu", None, 3, 64, 32, 2, 44),                     # exactly half a tile per pair
    ("favor_plus", "circulant_string", 2, 37, 32, 2, 200),
    ("favor_plus", "rope", 1, 33, 32, 2, 128),            # single pair per head: second tile half empty
]


@pytest.mark.parametrize("a,r,b,n,dim,heads,m", ORACLE_CASES)
def test_attention_matches_oracle(a, r, b, n, dim, heads, m):
    from erv_b200 import ATTENTION_REGISTRY, RPE_REGISTRY
    from oracle import erv_oracle as O
    torch.manual_seed(b * 131 + n * 7 + dim)
    kw = {"num_features": m} if m else {}
    attn = ATTENTION_REGISTRY[a](dim=dim, heads=heads, dropout=0.0, **kw)
    rpe = RPE_REGISTRY[r](num_patches=n, dim=dim, heads=heads) if r else None
    rpe_sd = {}
    if rpe is not None:
        with torch.no_grad():
            for p in rpe.parameters():
                p.normal_(0.0, 0.3)
        rpe_sd = {k: v.clone() for k, v in rpe.state_dict().items()}
    x = torch.randn(b, n, dim)
    w = torch.randn(b, n, dim)
    # oracle (CPU, fp32)
    params = {k: v.clone() for k, v in attn.state_dict().items()}
    leaves = [params[k].requires_grad_(True) for k in ("qkv.weight", "proj.weight", "proj.bias")]
    for k in ("rel_pos_bias", "circulant_coeffs"):
        if k in rpe_sd:
            leaves.append(rpe_sd[k].requires_grad_(True))
    if r == "rope":
        rpe_sd["cos"], rpe_sd["sin"] = O.rope_tables(n, dim // heads)
    xo = x.clone().requires_grad_(True)
    want = O.attention_forward(xo, params, heads, ATTN_KIND[a], RPE_KIND[r or "none"], rpe_sd, route="dense")
    (want * w).sum().backward()
    # CUDA path
    attn = attn.to(DEV).eval()
    rpe = rpe.to(DEV) if rpe is not None else None
    xg = x.to(DEV).requires_grad_(True)
    got = attn(xg, rpe=rpe)
    (got * w.to(DEV)).sum().backward()
    assert_close(got, want, TOL_F32, "out")
    assert_close(xg.grad, xo.grad, TOL_F32, "dx")
    assert_close(attn.qkv.weight.grad, leaves[0].grad, TOL_F32, "d qkv.weight")
    if rpe is not None:
        for p, l in zip(rpe.parameters(), leaves[3:]):
            assert_close(p.grad, l.grad, TOL_F32, "d rpe parameter")


@pytest.mark.parametrize("b,n,m,dtype", [(7, 65, 256, torch.float32), (3, 197, 44, torch.float32), (2, 300, 200, torch.float32),
                                         (3, 197, 44, torch.bfloat16)])
def test_linear_backward_without_saved_state(monkeypatch, b, n, m, dtype):
    """The tensor-core backward gives the same gradients whether it reads the forward's saved [S|z] or rebuilds it (short
    sequences: erv_linattn_pipe / tc2; N > 65: erv_linattn_tc, where the saved state replaces the K1 sweep)."""
    from erv_b200 import FAVORPlusAttention, ops
    torch.manual_seed(5)
    attn = FAVORPlusAttention(32, 2, num_features=m).to(DEV)
    qkv = torch.randn(b, n, 96, device=DEV).to(dtype)
    w = torch.randn(b, n, 32, device=DEV).to(dtype)
    grads = []
    for save in (True, False):
        monkeypatch.setattr(ops, "SAVE_KV_STATE", save)
        q = qkv.clone().requires_grad_(True)
        (ops.linear_attention(q, attn.omega, 2, ops.FEAT_FAVOR) * w).sum().backward()
        grads.append(q.grad.float())
    assert rel_l2(grads[0], grads[1]) < (1e-5 if dtype == torch.float32 else 2e-2)


def test_softmax_mask_and_return_attention():
    from erv_b200 import SoftmaxAttention
    from oracle import erv_oracle as O
    torch.manual_seed(3)
    attn = SoftmaxAttention(32, 2).eval()
    x = torch.randn(3, 70, 32)
    mask = (torch.rand(3, 70, 70) > 0.3).int()
    mask[:, torch.arange(70), torch.arange(70)] = 1
    params = attn.state_dict()
    qkv = torch.nn.functional.linear(x, params["qkv.weight"])
    q, k, v = O.split_qkv(qkv, 2)
    o_ref, p_ref = O.softmax_attention_core(q, k, v, None, mask=mask, return_attention=True)
    want = torch.nn.functional.linear(o_ref.transpose(1, 2).reshape(3, 70, 32), params["proj.weight"], params["proj.bias"])
    attn = attn.to(DEV)
    got, p = attn(x.to(DEV), mask=mask.to(DEV), return_attention=True)
    assert p.shape == (3, 2, 70, 70)
    assert rel_l2(got, want) < TOL_F32 and rel_l2(p, p_ref) < TOL_F32
    got4, _ = attn(x.to(DEV), mask=mask.to(DEV).unsqueeze(1), return_attention=True)
    assert torch.equal(got4, got)


def test_softmax_dropout_statistics_and_determinism():
    from erv_b200 import SoftmaxAttention, ops
    torch.manual_seed(0)
    attn = SoftmaxAttention(32, 2, dropout=0.25).to(DEV).train()
    x = torch.randn(4, 65, 32, device=DEV)
    qkv = attn.qkv(x)
    _, p = ops.softmax_attention(qkv, 2, dropout_p=0.25, seed=123, want_attn=True)
    _, p0 = ops.softmax_attention(qkv, 2, dropout_p=0.0, seed=123, want_attn=True)
    kept = (p != 0).float().mean().item()
    assert abs(kept - 0.75) < 0.02
    nz = p != 0
    assert torch.allclose(p[nz], p0[nz] / 0.75, rtol=1e-5, atol=1e-7)
    o1, _ = ops.softmax_attention(qkv, 2, dropout_p=0.25, seed=123)
    o2, _ = ops.softmax_attention(qkv, 2, dropout_p=0.25, seed=123)
    assert torch.equal(o1, o2)
    # backward uses the same mask: gradient equals autograd through the dumped probabilities
    qkv2 = qkv.detach().clone().requires_grad_(True)
    o, _ = ops.softmax_attention(qkv2, 2, dropout_p=0.25, seed=123)
    w = torch.randn_like(o)
    (o * w).sum().backward()
    qkv3 = qkv.detach().clone().requires_grad_(True)
    t = qkv3.reshape(4, 65, 3, 2, 16).permute(2, 0, 3, 1, 4)
    s = (t[0] @ t[1].transpose(-2, -1)) * 16 ** -0.5
    keep = (p != 0).float() / 0.75
    o_t = ((s.softmax(-1) * keep) @ t[2]).transpose(1, 2).reshape(4, 65, 32)
    (o_t * w).sum().backward()
    assert rel_l2(o, o_t) < TOL_F32 and rel_l2(qkv2.grad, qkv3.grad) < TOL_F32


def test_linear_path_properties_at_full_size():
    """BASELINE config-2 shape at B=1024 (too slow for the oracle): size-independent properties, NOT values (values are
    pinned at B <= 16 by the golden / oracle cases above, and (a) ties the big batch to the small one bit for bit).
    (a) batch independence: rows of a big batch equal the same rows run alone;
    (b) linearity in v: out(q,k,a*v1+b*v2) = a*out(v1)+b*out(v2);
    (c) a constant v = c gives c * den/(den + 1e-6): equal across the head's columns and in (0, c]."""
    from erv_b200 import FAVORPlusAttention, ops
    torch.manual_seed(1)
    attn = FAVORPlusAttention(32, 2, num_features=256).to(DEV)
    qkv = torch.randn(1024, 65, 96, device=DEV)
    out = ops.linear_attention(qkv, attn.omega, 2, ops.FEAT_FAVOR)
    sub = ops.linear_attention(qkv[500:508].contiguous(), attn.omega, 2, ops.FEAT_FAVOR)
    assert torch.equal(out[500:508], sub)
    q2 = qkv.clone().view(1024, 65, 3, 32)
    v1, v2 = q2[:, :, 2].clone(), torch.randn(1024, 65, 32, device=DEV)
    q2[:, :, 2] = v2
    out2 = ops.linear_attention(q2.view(1024, 65, 96), attn.omega, 2, ops.FEAT_FAVOR)
    q2[:, :, 2] = 0.5 * v1 - 2.0 * v2
    out3 = ops.linear_attention(q2.view(1024, 65, 96), attn.omega, 2, ops.FEAT_FAVOR)
    assert rel_l2(out3, 0.5 * out - 2.0 * out2) < TOL_F32
    q2[:, :, 2] = 3.0
    outc = ops.linear_attention(q2.view(1024, 65, 96), attn.omega, 2, ops.FEAT_FAVOR).view(1024, 65, 2, 16)
    assert (outc > 0).all() and (outc <= 3.0 * (1 + 1e-6)).all()
    assert (outc - outc[..., :1]).abs().max() < 1e-5


def test_edge_cases():
    from erv_b200 import (CirculantStringRPE, FAVORPlusAttention, KERPLEPositionalEncoding, ReLUAttention,
                          SoftmaxAttention)
    for n in (1, 2, 5):  # single token (CLS only), tiny sequences
        x = torch.randn(2, n, 32, device=DEV, requires_grad=True)
        for cls in (SoftmaxAttention, FAVORPlusAttention, ReLUAttention):
            attn = cls(32, 2).to(DEV)
            out = attn(x)
            out.sum().backward()
            assert torch.isfinite(out).all() and torch.isfinite(x.grad).all()
        rpe = KERPLEPositionalEncoding(n, 32, 2).to(DEV)
        assert torch.isfinite(ReLUAttention(32, 2).to(DEV)(x, rpe=rpe)).all()
    x = torch.randn(2, 1, 32, device=DEV)
    only_cls = CirculantStringRPE(1, 32, 2).to(DEV)
    assert torch.isfinite(FAVORPlusAttention(32, 2).to(DEV)(x, rpe=only_cls)).all()
    for scale in (10.0, 0.01):  # test_performer.py:177-196
        y = FAVORPlusAttention(32, 2).to(DEV)(torch.randn(2, 17, 32, device=DEV) * scale)
        assert torch.isfinite(y).all()
    with pytest.raises(NotImplementedError):
        FAVORPlusAttention(32, 2).to(DEV)(torch.randn(1, 4, 32, device=DEV), return_attention=True)
    with pytest.raises(NotImplementedError):  # head_dim 12 has no kernel instance: loud failure, no fallback
        FAVORPlusAttention(24, 2).to(DEV)(torch.randn(1, 4, 24, device=DEV))


def test_kerple_apply_rpe_fft_public_helper():
    from erv_b200 import KERPLEPositionalEncoding
    from oracle import erv_oracle as O
    torch.manual_seed(2)
    rpe = KERPLEPositionalEncoding(9, 32, 2)
    with torch.no_grad():
        rpe.rel_pos_bias.normal_(0, 0.3)
    kf, v = torch.rand(2, 2, 9, 6), torch.randn(2, 2, 9, 16)
    d1, d2 = O.kerple_d1_d2(kf, v, rpe.rel_pos_bias.detach())
    rpe = rpe.to(DEV)
    g1 = rpe.apply_rpe_fft(kf.to(DEV), v.to(DEV))
    g2 = rpe.apply_rpe_fft(kf.to(DEV))
    assert g1.shape == (2, 2, 9, 6, 16) and g2.shape == (2, 2, 9, 6)
    assert rel_l2(g1, d1) < TOL_F32 and rel_l2(g2, d2) < TOL_F32


def test_training_steps_are_stable():
    """10 Adam steps, loss stays finite and < 100 (test_kerple.py:380-411)."""
    from erv_b200 import MNIST_CONFIG, create_model
    torch.manual_seed(0)
    for name in ("performer_favor_most_general", "performer_relu_circulant", "baseline_rope"):
        model = create_model(name, MNIST_CONFIG).to(DEV).train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        img, lab = torch.randn(16, 1, 28, 28, device=DEV), torch.randint(0, 10, (16,), device=DEV)
        for _ in range(10):
            loss = torch.nn.functional.cross_entropy(model(img), lab)
            opt.zero_grad()
            loss.backward()
            opt.step()
        assert torch.isfinite(loss) and float(loss) < 100


def test_softmax_dropout_draws_new_masks_on_graph_replay():
    """The attention-dropout seed is device resident: a captured CUDA graph must not replay the same mask."""
    from erv_b200 import SoftmaxAttention, ops
    torch.manual_seed(0)
    attn = SoftmaxAttention(32, 2, dropout=0.3).to(DEV).train()
    x = torch.randn(4, 17, 32, device=DEV)
    with torch.no_grad():
        qkv = attn.qkv(x)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            attn.core(qkv, x.shape, None)  # warm-up: creates the per-device seed state outside the capture
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = attn.core(qkv, x.shape, None)
        g.replay()
        a = out.clone()
        g.replay()
        b = out.clone()
    assert torch.isfinite(a).all() and not torch.equal(a, b)
    # a fixed device seed reproduces the mask, forward and backward
    seed = torch.tensor([77], dtype=torch.int64, device=DEV)
    o1, _ = ops.softmax_attention(qkv, 2, dropout_p=0.3, seed=seed)
    o2, _ = ops.softmax_attention(qkv, 2, dropout_p=0.3, seed=seed)
    assert torch.equal(o1, o2)


# ---- KERPLE by FFT (erv_kerple_fft.cu): the reference's own route (kerple.py:252-270 -> fft_utils.py:142-170) ------------
@pytest.fixture
def kerple_fft_forced():
    from erv_b200 import _capi
    lib = _capi.load()
    lib.erv_kerple_set_fft(1)
    yield
    lib.erv_kerple_set_fft(-1)


@pytest.mark.parametrize("fname", [f for f in golden_files("attn_") if "most_general" in f])
def test_kerple_fft_route_matches_reference(fname, kerple_fft_forced):
    """Golden KERPLE cases with the forward forced onto the FFT route (the backward stays on the tile route and consumes
    the FFT route's saved output / normaliser), fp32 and bf16 autocast."""
    _, a, r = parse_attn_case(fname)
    g = load_golden(fname)
    attn, rpe = _build(g, a, r)
    x = g["x"].to(DEV).requires_grad_(True)
    out = attn(x, rpe=rpe)
    assert_close(out, g["out"], TOL_F32, "out")
    (out * g["cotangent"].to(DEV)).sum().backward()
    assert_close(x.grad, g["dx"], TOL_F32, "dx")
    for k, p in rpe.named_parameters():
        assert_close(p.grad, g["grad.rpe." + k], TOL_F32, k)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out16 = attn(x.detach(), rpe=rpe)
    assert out16.dtype == torch.bfloat16
    assert rel_l2(out16.float(), g["out"]) < TOL_BF16


@pytest.mark.parametrize("a,b,n,dim,heads,m", [
    ("favor_plus", 2, 258, 32, 2, 44),    # 257 patches: zero padding inside the thread's first register
    ("relu", 1, 1030, 32, 2, 33),         # odd feature count: the last transform carries one real column
    ("favor_plus", 1, 600, 64, 2, 16),    # head_dim 32
    ("favor_plus", 1, 4097, 32, 2, 44),   # config 5: all 4096 patches, circular length exactly 2*4096
    ("relu", 3, 2050, 32, 2, 256),        # M = 256, three batches, 2049 patches
])
def test_kerple_fft_route_matches_oracle(a, b, n, dim, heads, m, kerple_fft_forced):
    from erv_b200 import ATTENTION_REGISTRY, RPE_REGISTRY
    from oracle import erv_oracle as O
    torch.manual_seed(n + m)
    attn = ATTENTION_REGISTRY[a](dim=dim, heads=heads, dropout=0.0, num_features=m)
    rpe = RPE_REGISTRY["most_general"](num_patches=n, dim=dim, heads=heads)
    with torch.no_grad():
        rpe.rel_pos_bias.normal_(0.0, 0.3)
    x = torch.randn(b, n, dim)
    with torch.no_grad():
        want = O.attention_forward(x, dict(attn.state_dict()), heads, ATTN_KIND[a], "kerple", dict(rpe.state_dict()), route="dense")
        got = attn.to(DEV).eval()(x.to(DEV), rpe=rpe.to(DEV))
    assert_close(got, want, TOL_F32, "out")


def test_kerple_default_dispatch_with_many_pairs_takes_the_fft_route():
    """16 (batch, head) pairs at 2059 patches: the default dispatch (no forcing) sends the forward to the FFT route (three chunks
    of feature pairs per CTA column) and the backward to the tiles; outputs and gradients against the oracle's dense route."""
    from erv_b200 import ATTENTION_REGISTRY, RPE_REGISTRY, _capi
    from oracle import erv_oracle as O
    lib = _capi.load()
    b, n, dim, heads, m = 8, 2060, 32, 2, 44
    assert lib.erv_kerple_attention_workspace(b, n, heads, dim // heads, m, 0) > lib.erv_kerple_attention_workspace(2, n, heads, dim // heads, m, 0) * 4
    torch.manual_seed(11)
    attn = ATTENTION_REGISTRY["favor_plus"](dim=dim, heads=heads, dropout=0.0, num_features=m)
    rpe = RPE_REGISTRY["most_general"](num_patches=n, dim=dim, heads=heads)
    with torch.no_grad():
        rpe.rel_pos_bias.normal_(0.0, 0.3)
    x, w = torch.randn(b, n, dim), torch.randn(b, n, dim)
    params = {k: v.clone() for k, v in attn.state_dict().items()}
    bias = rpe.rel_pos_bias.detach().clone().requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    want = O.attention_forward(xo, params, heads, "favor", "kerple", {"rel_pos_bias": bias}, route="dense")
    (want * w).sum().backward()
    attn, rpe = attn.to(DEV).eval(), rpe.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    got = attn(xg, rpe=rpe)
    (got * w.to(DEV)).sum().backward()
    assert_close(got, want, TOL_F32, "out")
    assert_close(xg.grad, xo.grad, TOL_F32, "dx")
    assert_close(rpe.rel_pos_bias.grad, bias.grad, TOL_F32, "d rel_pos_bias")
