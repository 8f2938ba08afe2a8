"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference) on seeded inputs.  Run in the build container only:

    python oracle/make_golden.py

The reference cannot travel to the GPU box, the vectors can.  Each file holds the inputs, the
reference module's state_dict, its output and its autograd gradients, in float32, plus the
output/input-gradient under torch.autocast(bfloat16) for the bf16 tolerance tests.
"""
import copy
import os
import re
import sys

import numpy as np
import torch

REF = os.environ.get("ERV_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from models.attention import ATTENTION_REGISTRY  # noqa: E402
from models.rpe import RPE_REGISTRY  # noqa: E402
from models.rpe.fft_utils import fft_toeplitz_matmul  # noqa: E402
from models import create_model  # noqa: E402
from configs import MNIST_CONFIG, CIFAR10_CONFIG  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)

ATTN = ["softmax", "favor_plus", "relu"]
RPES = [None, "rope", "circulant_string", "most_general"]
# name: (B, N, dim, heads, num_features)
SHAPES = {
    "a": (2, 17, 32, 2, None),    # reference default dims, MNIST/CIFAR p8 token count, M=44
    "b": (3, 26, 64, 2, 64),      # Dh=32
    "c": (1, 65, 64, 1, 72),      # Dh=64 (one head), CIFAR p4 token count
    # shapes that dispatch to the tcgen05 kernels (Dh = 16): BASELINE config 2 (N=65, M=256: two pairs per tile) and
    # config 4 (N=197, M=44: one pair per 128-token tile, two token tiles); fp32 and bf16-autocast arms
    "d": (2, 65, 32, 2, 256),
    "e": (2, 197, 32, 2, 44),
}
ONLY = [a for a in sys.argv[1:] if a in SHAPES]  # e.g. `python oracle/make_golden.py d e` adds shapes without touching the rest


def np32(t):
    return t.detach().to(torch.float32).numpy().copy()  # copy: .grad buffers are accumulated into later


def build(attn_name, rpe_name, n, dim, heads, m, seed):
    torch.manual_seed(seed)
    kw = {}
    if attn_name != "softmax" and m is not None:
        kw["num_features"] = m
    attn = ATTENTION_REGISTRY[attn_name](dim=dim, heads=heads, dropout=0.0, **kw)
    rpe = None
    if rpe_name is not None:
        rpe = RPE_REGISTRY[rpe_name](num_patches=n, dim=dim, heads=heads)
        with torch.no_grad():  # make the RPE matter (init is N(0, 0.02^2) / N(0, 0.01^2))
            if rpe_name == "most_general":
                rpe.rel_pos_bias.normal_(0.0, 0.4)
            if rpe_name == "circulant_string":
                rpe.circulant_coeffs.normal_(0.0, 0.3)
    return attn, rpe


def attention_cases():
    seed = 1000
    for sname, (b, n, dim, heads, m) in SHAPES.items():
        for a in ATTN:
            for r in RPES:
                if a == "softmax" and r == "most_general":
                    continue
                seed += 1
                if ONLY and sname not in ONLY:
                    continue
                # ReLU features are not differentiable where a projection crosses zero: a pre-activation within rounding
                # distance of 0 makes the REFERENCE's own gradient jump under a 1e-6 perturbation of its input, and no
                # implementation can be asked to reproduce that.  Shapes added in round 2 (d, e) are screened: the reference's
                # dx must move by < 1e-4 (rel-L2) under 2e-6 relative perturbations (smooth cases move by ~1e-5), else the case is re-seeded (recorded in the file).
                case_seed, reseeds = seed, 0
                while True:
                    attn, rpe = build(a, r, n, dim, heads, m, case_seed)
                    attn.eval()
                    x = torch.randn(b, n, dim, requires_grad=True)
                    w = torch.randn(b, n, dim)  # cotangent
                    out = attn(x, rpe=rpe)
                    (out * w).sum().backward()
                    if sname not in ("d", "e"):
                        break
                    drift = 0.0
                    for _ in range(3):  # three draws of a 2e-6 relative perturbation (what a different rounding of the rotation does)
                        x_p = (x.detach() * (1.0 + 2e-6 * torch.randn(b, n, dim))).requires_grad_(True)
                        (attn(x_p, rpe=rpe) * w).sum().backward()
                        drift = max(drift, float((x_p.grad - x.grad).norm() / x.grad.norm()))
                    for p_ in list(attn.parameters()) + (list(rpe.parameters()) if rpe is not None else []):
                        p_.grad = None
                    if drift < 1e-4:
                        x.grad = None
                        out = attn(x, rpe=rpe)
                        (out * w).sum().backward()
                        break
                    reseeds += 1
                    print(f"  {sname} {a} {r}: reference dx drifts {drift:.1e} under 2e-6 input perturbations (kink); re-seeding")
                    case_seed += 100000
                rec = {"x": np32(x), "cotangent": np32(w), "out": np32(out), "dx": np32(x.grad),
                       "heads": np.int64(heads), "seed": np.int64(case_seed)}
                for k, v in attn.state_dict().items():
                    rec["attn." + k] = v.numpy()
                for k, p in attn.named_parameters():
                    rec["grad.attn." + k] = np32(p.grad)
                if rpe is not None:
                    for k, v in rpe.state_dict().items():
                        rec["rpe." + k] = v.numpy()
                    for k, p in rpe.named_parameters():
                        rec["grad.rpe." + k] = np32(p.grad)
                # bf16 autocast arm (fresh grads)
                x2 = x.detach().clone().requires_grad_(True)
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    out16 = attn(x2, rpe=rpe)
                (out16.float() * w).sum().backward()
                rec["out_bf16"] = np32(out16)
                rec["dx_bf16"] = np32(x2.grad)
                np.savez_compressed(os.path.join(OUT, f"attn_{sname}_{a}_{r or 'none'}.npz"), **rec)
                print("attn", sname, a, r, tuple(out.shape))


def unit_cases():
    torch.manual_seed(7)
    rec = {}
    # Toeplitz product (fft_utils.py), incl. the docstring example at fft_utils.py:276-281
    c = torch.randn(2 * 8 - 1)
    x = torch.randn(8, 4)
    rec["toep.c"], rec["toep.x"], rec["toep.y"] = np32(c), np32(x), np32(fft_toeplitz_matmul(c, x))
    cb = torch.randn(2, 3, 2 * 33 - 1)
    xb = torch.randn(2, 3, 33, 5)
    rec["toepb.c"], rec["toepb.x"], rec["toepb.y"] = np32(cb), np32(xb), np32(fft_toeplitz_matmul(cb, xb))
    # feature maps
    fav = ATTENTION_REGISTRY["favor_plus"](dim=32, heads=2)
    rel = ATTENTION_REGISTRY["relu"](dim=32, heads=2)
    q = torch.randn(2, 2, 9, 16)
    rec["feat.x"] = np32(q)
    rec["feat.omega_favor"], rec["feat.phi_favor"] = fav.omega.numpy(), np32(fav._compute_phi_positive(q, fav.omega))
    rec["feat.omega_relu"], rec["feat.phi_relu"] = rel.omega.numpy(), np32(rel._compute_relu_features(q, rel.omega))
    # rotations
    rope = RPE_REGISTRY["rope"](num_patches=10, dim=32, heads=2)
    k = torch.randn(2, 2, 9, 16)
    qr, kr = rope.apply_rotary_emb(q, k)
    rec["rope.k"], rec["rope.q_out"], rec["rope.k_out"] = np32(k), np32(qr), np32(kr)
    rec["rope.cos"], rec["rope.sin"] = rope.cos_cached.numpy(), rope.sin_cached.numpy()
    circ = RPE_REGISTRY["circulant_string"](num_patches=10, dim=32, heads=2)
    with torch.no_grad():
        circ.circulant_coeffs.normal_(0.0, 0.3)
    q10, k10 = torch.randn(2, 2, 10, 16), torch.randn(2, 2, 10, 16)
    qc, kc = circ.apply_circulant_string(q10, k10)
    rec["circ.coeffs"], rec["circ.pos"] = np32(circ.circulant_coeffs), circ.patch_positions.numpy()
    rec["circ.q"], rec["circ.k"], rec["circ.q_out"], rec["circ.k_out"] = np32(q10), np32(k10), np32(qc), np32(kc)
    ev = circ.get_eigenvalues()
    rec["circ.eig_real"], rec["circ.eig_imag"] = np32(ev.real), np32(ev.imag)
    np.savez_compressed(os.path.join(OUT, "units.npz"), **rec)
    print("units done")


# model-level gradients kept in the fixtures (the rest only bloat the files)
GRAD_KEYS = re.compile(r"cls_token|transformer_blocks\.0\.attention|transformer_blocks\.\d\.rpe|mlp_head\.1")


def model_cases():
    from models.factory import MODEL_VARIANTS
    names = [n for n in MODEL_VARIANTS if n not in ("performer", "vit", "baseline_most_general")]
    for i, name in enumerate(names):
        torch.manual_seed(2000 + i)
        cfg = copy.deepcopy(MNIST_CONFIG)
        model = create_model(name, cfg, dropout=0.0)
        model.eval()
        with torch.no_grad():
            for k, p in model.named_parameters():
                if k.endswith("rel_pos_bias"):
                    p.normal_(0.0, 0.4)
                if k.endswith("circulant_coeffs"):
                    p.normal_(0.0, 0.3)
        img = torch.randn(4, 1, 28, 28)
        lab = torch.randint(0, 10, (4,))
        logits = model(img)
        loss = torch.nn.functional.cross_entropy(logits, lab)
        loss.backward()
        rec = {"images": np32(img), "labels": lab.numpy(), "logits": np32(logits), "loss": np32(loss)}
        for k, v in model.state_dict().items():
            rec["sd." + k] = v.numpy()
        for k, p in model.named_parameters():
            if GRAD_KEYS.search(k):
                rec["grad." + k] = np32(p.grad)
        np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **rec)
        print("model", name, float(loss.detach()))
    # one CIFAR-shaped case: BASELINE config 2 (performer_favor, p4, M=256), tiny batch
    torch.manual_seed(3000)
    model = create_model("performer_favor", copy.deepcopy(CIFAR10_CONFIG),
                         attention_config={"num_features": 256}, patch_size=4, dropout=0.0)
    model.eval()
    img = torch.randn(2, 3, 32, 32)
    lab = torch.randint(0, 10, (2,))
    logits = model(img)
    loss = torch.nn.functional.cross_entropy(logits, lab)
    loss.backward()
    rec = {"images": np32(img), "labels": lab.numpy(), "logits": np32(logits), "loss": np32(loss)}
    for k, v in model.state_dict().items():
        rec["sd." + k] = v.numpy()
    for k, p in model.named_parameters():
        if GRAD_KEYS.search(k):
            rec["grad." + k] = np32(p.grad)
    np.savez_compressed(os.path.join(OUT, "model_cifar_performer_favor_m256.npz"), **rec)
    print("model cifar m256", float(loss.detach()))


if __name__ == "__main__":
    attention_cases()
    if not ONLY:
        unit_cases()
        model_cases()
