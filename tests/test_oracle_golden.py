"""Pins oracle/erv_oracle.py to outputs of the unmodified reference (tests/golden/, produced by
oracle/make_golden.py).  CPU only.  Tolerance: fp32 rel-L2 <= 2e-5 (both sides are fp32 eager
torch; only op ordering differs)."""
import pytest
import torch

from conftest import ATTN_KIND, RPE_KIND, golden_files, load_golden, parse_attn_case, rel_l2
from oracle import erv_oracle as O

TOL = 2e-5


def _attn_inputs(g, a, r):
    params = {k[5:]: v for k, v in g.items() if k.startswith("attn.")}
    rpe = {k[4:]: v for k, v in g.items() if k.startswith("rpe.")}
    heads = int(g["heads"])
    if r == "rope":
        n, dh = g["x"].shape[1], g["x"].shape[2] // heads
        rpe["cos"], rpe["sin"] = O.rope_tables(n, dh)
    return params, rpe, heads


@pytest.mark.parametrize("fname", golden_files("attn_"))
@pytest.mark.parametrize("route", ["fft", "dense"])
def test_attention_matches_reference(fname, route):
    _, a, r = parse_attn_case(fname)
    if route == "dense" and r != "most_general":
        pytest.skip("dense route only differs for KERPLE")
    g = load_golden(fname)
    params, rpe, heads = _attn_inputs(g, a, r)
    leaves = {}
    for k in ("qkv.weight", "proj.weight", "proj.bias"):
        leaves["attn." + k] = params[k] = params[k].clone().requires_grad_(True)
    for k in ("rel_pos_bias", "circulant_coeffs"):
        if k in rpe:
            leaves["rpe." + k] = rpe[k] = rpe[k].clone().requires_grad_(True)
    x = g["x"].clone().requires_grad_(True)
    out = O.attention_forward(x, params, heads, ATTN_KIND[a], RPE_KIND[r or "none"], rpe, route=route)
    assert rel_l2(out, g["out"]) < TOL
    (out * g["cotangent"]).sum().backward()
    assert rel_l2(x.grad, g["dx"]) < 5 * TOL
    for k, t in leaves.items():
        assert rel_l2(t.grad, g["grad." + k]) < 5 * TOL, k


def test_units_match_reference():
    g = load_golden("units.npz")
    assert rel_l2(O.toeplitz_matmul(g["toep.c"], g["toep.x"]), g["toep.y"]) < TOL
    assert rel_l2(O.toeplitz_matmul(g["toepb.c"], g["toepb.x"]), g["toepb.y"]) < TOL
    n = g["toep.x"].shape[0]
    assert rel_l2(O.toeplitz_dense(g["toep.c"], n) @ g["toep.x"], g["toep.y"]) < TOL
    assert rel_l2(O.favor_features(g["feat.x"], g["feat.omega_favor"]), g["feat.phi_favor"]) < TOL
    assert rel_l2(O.relu_features(g["feat.x"], g["feat.omega_relu"]), g["feat.phi_relu"]) < TOL
    cos, sin = O.rope_tables(10, 16)
    assert torch.equal(cos, g["rope.cos"]) and torch.equal(sin, g["rope.sin"])
    assert rel_l2(O.rope_rotate(g["feat.x"], cos, sin), g["rope.q_out"]) < TOL
    assert rel_l2(O.rope_rotate(g["rope.k"], cos, sin), g["rope.k_out"]) < TOL
    assert torch.equal(O.circulant_positions(10), g["circ.pos"])
    assert rel_l2(O.circulant_rotate(g["circ.q"], g["circ.coeffs"], g["circ.pos"]), g["circ.q_out"]) < TOL
    assert rel_l2(O.circulant_rotate(g["circ.k"], g["circ.coeffs"], g["circ.pos"]), g["circ.k_out"]) < TOL
    ev = O.circulant_eigenvalues(g["circ.coeffs"])
    assert ev.real.abs().max() < 1e-6 and rel_l2(ev.imag, g["circ.eig_imag"]) < TOL


def test_toeplitz_docstring_example():
    # fft_utils.py:276-281
    c = torch.tensor([4.0, 3.0, 2.0, 1.0, 2.0, 3.0, 4.0])
    t = O.toeplitz_dense(c, 4)
    want = torch.tensor([[1., 2, 3, 4], [2, 1, 2, 3], [3, 2, 1, 2], [4, 3, 2, 1]])
    assert torch.equal(t, want)


@pytest.mark.parametrize("fname", golden_files("model_"))
def test_model_matches_reference(fname):
    g = load_golden(fname)
    name = fname[len("model_"):-4]
    if name.startswith("cifar_"):
        name, cfg = "performer_favor", dict(patch_size=4, heads=2, depth=3, dim=32)
    else:
        cfg = dict(patch_size=7, heads=2, depth=3, dim=32)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    want = {k[5:]: v for k, v in g.items() if k.startswith("grad.")}
    for k in want:
        sd[k] = sd[k].clone().requires_grad_(True)
    logits = O.vit_forward(sd, g["images"], name, cfg)
    assert rel_l2(logits, g["logits"]) < TOL
    loss = torch.nn.functional.cross_entropy(logits, g["labels"])
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    for k, v in want.items():
        assert rel_l2(sd[k].grad, v) < 1e-4, k


def test_cpu_trainer_step_runs():
    cfg = dict(image_size=28, in_channels=1, patch_size=7, num_classes=10, dim=32, depth=3, heads=2, mlp_dim=64)
    for name in ("baseline", "performer_favor_most_general", "performer_relu_circulant"):
        tr = O.CpuTrainer(name, cfg, seed=0)
        l0 = tr.step(torch.randn(8, 1, 28, 28), torch.randint(0, 10, (8,)))
        assert l0 == l0 and l0 < 100
