"""CPU oracle for the efficient-rpe-vit attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``efficient-rpe-vit_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

What it is: a functional (no nn.Module) restatement, on CPU torch tensors, of what the
reference computes on the path SURVEY.md section 8 scopes.  All of the reference's arithmetic
is torch ATen (requirements.txt: torch>=2.0, unpinned; installed 2.11.0), so the oracle uses
the same primitive ops (einsum, exp, amax, torch.fft) and takes gradients with autograd,
exactly as the reference does (it has no custom backward anywhere).  Every function cites the
reference lines it follows.

Parity pin: ``oracle/make_golden.py`` imports the *unmodified* reference from
/root/reference, runs every attention x RPE combination plus the full models on seeded
inputs, and commits inputs/parameters/outputs/gradients under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks this file against those vectors (fp32 tolerance
1e-5), so the oracle is pinned to outputs of the reference itself.

Parameter tensors are passed in dicts that use the reference's ``state_dict`` key names
(``qkv.weight``, ``proj.weight``, ``proj.bias``, ``omega``, ``rel_pos_bias``,
``circulant_coeffs``, ``patch_positions``), so golden state dicts plug in directly.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

EPS = 1e-6  # favor_plus.py:260 / relu.py:258: added to the denominator, not clamped


# --------------------------------------------------------------------------------------
# RoPE  (models/rpe/rope.py)
# --------------------------------------------------------------------------------------
def rope_tables(num_patches: int, head_dim: int, theta: float = 10000.0, dtype=torch.float32):
    """cos/sin caches [num_patches, head_dim/2]; rope.py:53-68."""
    m = torch.arange(0, head_dim, 2).float() / head_dim
    freqs = 1.0 / (theta ** m)
    ang = torch.arange(num_patches).float()[:, None] * freqs[None, :]
    return torch.cos(ang).to(dtype), torch.sin(ang).to(dtype)


def rope_rotate(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, positions=None):
    """Interleaved-pair rotation of x [B,H,N,Dh]; rope.py:95-135.  CLS sits at position 0."""
    n = x.shape[2]
    idx = torch.arange(n) if positions is None else positions
    c, s = cos[idx], sin[idx]  # [N, Dh/2]
    xe, xo = x[..., 0::2], x[..., 1::2]
    re = xe * c - xo * s
    ro = xe * s + xo * c
    return torch.stack([re, ro], dim=-1).reshape(x.shape)


# --------------------------------------------------------------------------------------
# Circulant-STRING  (models/rpe/circulant_string.py)
# --------------------------------------------------------------------------------------
def circulant_positions(num_patches: int, coord_dim: int = 2) -> torch.Tensor:
    """[x, y] integer grid, row-major, for the N-1 patch tokens; circulant_string.py:160-205."""
    n = num_patches - 1
    if n <= 0:
        return torch.zeros(0, coord_dim)
    side = int(math.sqrt(n))
    if side * side != n:
        raise ValueError(f"num_patches - 1 = {n} must be a perfect square")
    ys, xs = torch.meshgrid(torch.arange(side, dtype=torch.float32),
                            torch.arange(side, dtype=torch.float32), indexing="ij")
    return torch.stack([xs.flatten(), ys.flatten()], dim=-1)


def circulant_eigenvalues(coeffs: torch.Tensor) -> torch.Tensor:
    """lambda(C - C^T) = FFT(c) - conj(FFT(c)); circulant_string.py:207-232."""
    lam = torch.fft.fft(coeffs, dim=-1)
    return lam - torch.conj(lam)


def circulant_rotate_patches(x: torch.Tensor, coeffs: torch.Tensor, positions: torch.Tensor):
    """x' = Re IFFT(exp(mu) * FFT(x)), mu[h,n,:] = sum_k pos[n,k] * lambda[h,k,:].

    x is [B,H,Np,Dh] WITHOUT the CLS row; circulant_string.py:234-295.  Like the reference,
    the transform runs in complex64 for float32 input (complex128 for float64).
    """
    _, h, n, d = x.shape
    eig = circulant_eigenvalues(coeffs)  # [H, K, Dh]
    mu = (positions.view(1, 1, n, -1, 1) * eig.view(1, h, 1, -1, d)).sum(dim=-2)
    ctype = torch.complex128 if x.dtype == torch.float64 else torch.complex64
    xf = torch.fft.fft(x.to(ctype), dim=-1)
    y = torch.fft.ifft(torch.exp(mu) * xf, dim=-1).real
    return y.to(x.dtype)


def circulant_rotate(x: torch.Tensor, coeffs: torch.Tensor, positions: torch.Tensor):
    """CLS (index 0) passes through, patches are rotated; circulant_string.py:297-341."""
    if x.shape[2] <= 1:
        return x
    rot = circulant_rotate_patches(x[:, :, 1:, :], coeffs, positions)
    return torch.cat([x[:, :, :1, :], rot], dim=2)


# --------------------------------------------------------------------------------------
# Toeplitz products  (models/rpe/fft_utils.py)
# --------------------------------------------------------------------------------------
def toeplitz_dense(c: torch.Tensor, n: int) -> torch.Tensor:
    """T[i,j] = c[j - i + n - 1]; fft_utils.py:261-292."""
    i = torch.arange(n)
    return c[(i[None, :] - i[:, None]) + (n - 1)]


def toeplitz_matmul_fft(c: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """y = T x through a length-(2n-1) circulant embedding; fft_utils.py:112-172.

    c: [2n-1] (shared) or [B, 2n-1]; x: [B, n, d].
    """
    if c.dim() == 1:
        c = c.unsqueeze(0).expand(x.shape[0], -1)
    n = (c.shape[1] + 1) // 2
    assert x.shape[1] == n
    col = torch.cat([c[:, n - 1:n], torch.flip(c[:, :n - 1], dims=[1]),
                     torch.flip(c[:, n:], dims=[1])], dim=1)
    cf = torch.fft.fft(col, dim=-1)
    xf = torch.fft.fft(F.pad(x, (0, 0, 0, n - 1)), dim=1)
    return torch.fft.ifft(cf.unsqueeze(-1) * xf, dim=1)[:, :n, :].real


def toeplitz_matmul(c: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """Shape dispatcher of fft_utils.py:17-84 (same accepted shapes, same errors)."""
    if c.dim() == 1:
        if x.dim() == 3:
            return toeplitz_matmul_fft(c, x)
        if x.dim() == 2:
            return toeplitz_matmul_fft(c, x.unsqueeze(0))[0]
        if x.dim() == 1:
            return toeplitz_matmul_fft(c, x.view(1, -1, 1))[0, :, 0]
        raise ValueError(f"x must have 1 or 2 dimensions. Got shape={tuple(x.shape)}")
    if c.dim() == 3:
        if x.dim() != 4:
            raise ValueError(f"When c has 3 dims, x must have 4 dims. Got x.shape={tuple(x.shape)}")
        assert c.shape[:2] == x.shape[:2]
        return torch.stack([toeplitz_matmul_fft(c[:, h], x[:, h]) for h in range(c.shape[1])], dim=1)
    raise ValueError(f"c must have 1 or 3 dimensions. Got shape={tuple(c.shape)}")


# --------------------------------------------------------------------------------------
# Random-feature maps  (favor_plus.py:112-140, relu.py:116-138)
# --------------------------------------------------------------------------------------
def favor_features(x: torch.Tensor, omega: torch.Tensor) -> torch.Tensor:
    proj = torch.einsum("bhnd,hdf->bhnf", x, omega)
    proj = proj - proj.amax(dim=-1, keepdim=True).detach()
    half_sq = (x ** 2).sum(dim=-1, keepdim=True) / 2.0
    return torch.exp(proj - half_sq) / math.sqrt(omega.shape[-1])


def relu_features(x: torch.Tensor, omega: torch.Tensor) -> torch.Tensor:
    return F.relu(torch.einsum("bhnd,hdf->bhnf", x, omega)) / math.sqrt(omega.shape[-1])


def orthogonal_omega(heads: int, head_dim: int, num_features: int) -> torch.Tensor:
    """QR-orthogonal blocks, truncated, times sqrt(Dh); favor_plus.py:83-110."""
    out = []
    for _ in range(heads):
        if num_features <= head_dim:
            q, _ = torch.linalg.qr(torch.randn(head_dim, num_features), mode="reduced")
            w = q * math.sqrt(head_dim)
        else:
            blocks = []
            for _ in range(math.ceil(num_features / head_dim)):
                q, _ = torch.linalg.qr(torch.randn(head_dim, head_dim), mode="reduced")
                blocks.append(q)
            w = torch.cat(blocks, dim=1)[:, :num_features] * math.sqrt(head_dim)
        out.append(w)
    return torch.stack(out, dim=0)


# --------------------------------------------------------------------------------------
# Attention cores.  q, k, v are [B,H,N,Dh] (views into the packed qkv buffer).
# --------------------------------------------------------------------------------------
def split_qkv(qkv: torch.Tensor, heads: int):
    """[B,N,3C] -> three [B,H,N,Dh] views; favor_plus.py:174-176, softmax.py:82-84."""
    b, n, c3 = qkv.shape
    dh = c3 // 3 // heads
    t = qkv.reshape(b, n, 3, heads, dh).permute(2, 0, 3, 1, 4)
    return t[0], t[1], t[2]


def _rotate_qk(q, k, rpe_kind: Optional[str], rpe: Dict[str, torch.Tensor]):
    if rpe_kind == "rope":
        return rope_rotate(q, rpe["cos"], rpe["sin"]), rope_rotate(k, rpe["cos"], rpe["sin"])
    if rpe_kind == "circulant":
        return (circulant_rotate(q, rpe["circulant_coeffs"], rpe["patch_positions"]),
                circulant_rotate(k, rpe["circulant_coeffs"], rpe["patch_positions"]))
    return q, k


def kerple_d1_d2(k_feat: torch.Tensor, v: torch.Tensor, bias: torch.Tensor):
    """D1[i] = sum_j c[j-i] phi(k_j)^T v_j and D2[i] = sum_j c[j-i] phi(k_j), c = exp(b).

    The reference's route (kerple.py:151-344): a materialised [B,H,N,M,Dh] outer product pushed
    through the FFT-Toeplitz product head by head.
    """
    b, h, n, m = k_feat.shape
    c = torch.exp(bias)
    a1 = torch.einsum("bhkf,bhkd->bhkfd", k_feat, v).reshape(b, h, n, m * v.shape[-1])
    d1 = torch.stack([toeplitz_matmul_fft(c[i], a1[:, i]) for i in range(h)], dim=1)
    d2 = torch.stack([toeplitz_matmul_fft(c[i], k_feat[:, i]) for i in range(h)], dim=1)
    return d1.reshape(b, h, n, m, v.shape[-1]), d2


def linear_attention_core(q, k, v, omega, kind: str, rpe_kind: Optional[str],
                          rpe: Optional[Dict[str, torch.Tensor]] = None, route: str = "fft"):
    """FAVOR+/ReLU linear attention on split heads -> [B,H,N,Dh].

    favor_plus.py:179-260 / relu.py:177-258.  ``route='dense'`` evaluates the KERPLE branch as
    Toeplitz-masked quadratic attention (identical to the FFT route).
    """
    dh = q.shape[-1]
    feat = favor_features if kind == "favor" else relu_features
    if rpe_kind == "kerple":
        q = q / torch.norm(q, p=2, dim=-1, keepdim=True)
        k = k / torch.norm(k, p=2, dim=-1, keepdim=True)
    else:
        q, k = _rotate_qk(q, k, rpe_kind, rpe or {})
        q = q * dh ** -0.25
        k = k * dh ** -0.25
    qf, kf = feat(q, omega), feat(k, omega)
    if rpe_kind == "kerple":
        if route == "fft":
            d1, d2 = kerple_d1_d2(kf, v, rpe["rel_pos_bias"])
            num = torch.einsum("bhnf,bhnfd->bhnd", qf, d1)
            den = torch.einsum("bhnf,bhnf->bhn", qf, d2)
        else:
            n = q.shape[2]
            t = torch.stack([toeplitz_dense(torch.exp(rpe["rel_pos_bias"][i]), n)
                             for i in range(q.shape[1])], dim=0)
            a = torch.einsum("bhif,bhjf->bhij", qf, kf) * t.unsqueeze(0)
            num = a @ v
            den = a.sum(dim=-1)
    else:
        kv = torch.einsum("bhnf,bhnd->bhfd", kf, v)
        num = torch.einsum("bhnf,bhfd->bhnd", qf, kv)
        den = torch.einsum("bhnf,bhf->bhn", qf, kf.sum(dim=2))
    return num / (den.unsqueeze(-1) + EPS)


def softmax_attention_core(q, k, v, rpe_kind: Optional[str], rpe=None, mask=None,
                           return_attention: bool = False, dropout: float = 0.0):
    """dropout(softmax(q k^T / sqrt(Dh) [+mask])) v; softmax.py:86-115."""
    if rpe_kind == "kerple":
        raise NotImplementedError("KERPLE RPE is designed specifically for kernelized attention")
    q, k = _rotate_qk(q, k, rpe_kind, rpe or {})
    s = (q @ k.transpose(-2, -1)) * q.shape[-1] ** -0.5
    if mask is not None:
        if mask.dim() == 3:
            mask = mask.unsqueeze(1)
        s = s.masked_fill(mask == 0, float("-inf"))
    p = F.dropout(s.softmax(dim=-1), dropout, training=dropout > 0)
    out = p @ v
    return (out, p) if return_attention else out


def attention_forward(x: torch.Tensor, params: Dict[str, torch.Tensor], heads: int, kind: str,
                      rpe_kind: Optional[str] = None, rpe: Optional[Dict[str, torch.Tensor]] = None,
                      mask=None, route: str = "fft", dropout: float = 0.0):
    """One attention module, x [B,N,C] -> [B,N,C]: qkv Linear, core, proj Linear, dropout (training only;
    softmax.py:112,120, favor_plus.py:265).

    kind in {'softmax','favor','relu'}; rpe_kind in {None,'rope','circulant','kerple'}.
    """
    b, n, c = x.shape
    qkv = F.linear(x, params["qkv.weight"], params.get("qkv.bias"))
    q, k, v = split_qkv(qkv, heads)
    if kind == "softmax":
        o = softmax_attention_core(q, k, v, rpe_kind, rpe, mask, dropout=dropout)
    else:
        o = linear_attention_core(q, k, v, params["omega"], kind, rpe_kind, rpe, route)
    o = o.transpose(1, 2).reshape(b, n, c)
    return F.dropout(F.linear(o, params["proj.weight"], params["proj.bias"]), dropout, training=dropout > 0)


# --------------------------------------------------------------------------------------
# Whole model (the caller of the hot path; needed for the images/sec CPU baseline)
# --------------------------------------------------------------------------------------
_ATTN_KIND = {"softmax": "softmax", "favor_plus": "favor", "relu": "relu"}
_RPE_KIND = {None: None, "rope": "rope", "circulant_string": "circulant", "most_general": "kerple"}

MODEL_VARIANTS = {
    "baseline": ("softmax", None),
    "baseline_circulant": ("softmax", "circulant_string"),
    "baseline_rope": ("softmax", "rope"),
    "performer_favor": ("favor_plus", None),
    "performer_favor_most_general": ("favor_plus", "most_general"),
    "performer_favor_circulant": ("favor_plus", "circulant_string"),
    "performer_favor_rope": ("favor_plus", "rope"),
    "performer_relu": ("relu", None),
    "performer_relu_most_general": ("relu", "most_general"),
    "performer_relu_circulant": ("relu", "circulant_string"),
    "performer_relu_rope": ("relu", "rope"),
}


def patchify(img: torch.Tensor, p: int) -> torch.Tensor:
    """[B,C,H,W] -> [B, (H/p)(W/p), C p p]; base_vit.py:174-198."""
    b, c, h, w = img.shape
    x = img.reshape(b, c, h // p, p, w // p, p).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(b, (h // p) * (w // p), c * p * p)


def vit_forward(sd: Dict[str, torch.Tensor], images: torch.Tensor, model_name: str, cfg: dict,
                route: str = "fft", dropout: float = 0.0) -> torch.Tensor:
    """Forward of the reference model from its state_dict (eval mode when dropout == 0); base_vit.py:200-233
    and unified_transformer.py:64-90.  cfg needs patch_size, heads, depth, dim (+ rope theta)."""
    attn_type, rpe_type = MODEL_VARIANTS[model_name]
    kind, rpe_kind = _ATTN_KIND[attn_type], _RPE_KIND[rpe_type]
    heads, dim = cfg["heads"], cfg["dim"]
    x = F.linear(patchify(images, cfg["patch_size"]), sd["patch_embedding.weight"], sd["patch_embedding.bias"])
    x = torch.cat([sd["cls_token"].expand(x.shape[0], -1, -1), x], dim=1) + sd["pos_embedding"]
    n = x.shape[1]
    for i in range(cfg["depth"]):
        pre = f"transformer_blocks.{i}."
        attn = {k[len(pre) + 10:]: v for k, v in sd.items() if k.startswith(pre + "attention.")}
        rpe = {k[len(pre) + 4:]: v for k, v in sd.items() if k.startswith(pre + "rpe.")}
        if rpe_kind == "rope":
            rpe["cos"], rpe["sin"] = rope_tables(n, dim // heads, cfg.get("theta", 10000.0), x.dtype)
        h = F.layer_norm(x, (dim,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
        x = x + attention_forward(h, attn, heads, kind, rpe_kind, rpe, route=route, dropout=dropout)
        h = F.layer_norm(x, (dim,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
        h = F.dropout(F.gelu(F.linear(h, sd[pre + "mlp.0.weight"], sd[pre + "mlp.0.bias"])), dropout, training=dropout > 0)
        x = x + F.dropout(F.linear(h, sd[pre + "mlp.3.weight"], sd[pre + "mlp.3.bias"]), dropout, training=dropout > 0)
    h = F.layer_norm(x[:, 0], (dim,), sd["mlp_head.0.weight"], sd["mlp_head.0.bias"])
    return F.linear(h, sd["mlp_head.1.weight"], sd["mlp_head.1.bias"])


_BUFFER_SUFFIXES = ("omega", "redraw_counter", "patch_positions")


def init_state(model_name: str, cfg: dict, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init state with the reference's key names, shapes and init laws
    (base_vit.py:153-172 xavier/zeros/ones, N(0,0.02) tokens; kerple.py:67-73; circulant_string.py:147-155).
    Values are NOT bit-equal to a reference construction (different RNG call order)."""
    g = torch.Generator().manual_seed(seed)
    attn_type, rpe_type = MODEL_VARIANTS[model_name]
    dim, heads, depth, mlp = cfg["dim"], cfg["heads"], cfg["depth"], cfg["mlp_dim"]
    dh = dim // heads
    n = (cfg["image_size"] // cfg["patch_size"]) ** 2 + 1
    pdim = cfg["in_channels"] * cfg["patch_size"] ** 2

    def xavier(o, i):
        a = math.sqrt(6.0 / (i + o))
        return (torch.rand(o, i, generator=g) * 2 - 1) * a

    sd = {"cls_token": torch.randn(1, 1, dim, generator=g) * 0.02,
          "pos_embedding": torch.randn(1, n, dim, generator=g) * 0.02,
          "patch_embedding.weight": xavier(dim, pdim), "patch_embedding.bias": torch.zeros(dim)}
    for i in range(depth):
        pre = f"transformer_blocks.{i}."
        sd[pre + "attention.qkv.weight"] = xavier(3 * dim, dim)
        sd[pre + "attention.proj.weight"] = xavier(dim, dim)
        sd[pre + "attention.proj.bias"] = torch.zeros(dim)
        if attn_type != "softmax":
            m = cfg.get("num_features") or int(dh * math.log(dh))
            torch.manual_seed(seed + 1000 + i)
            sd[pre + "attention.omega"] = orthogonal_omega(heads, dh, m)
        if rpe_type == "most_general":
            sd[pre + "rpe.rel_pos_bias"] = torch.randn(heads, 2 * n - 1, generator=g) * 0.02
        if rpe_type == "circulant_string":
            sd[pre + "rpe.circulant_coeffs"] = torch.randn(heads, 2, dh, generator=g) * 0.01
            sd[pre + "rpe.patch_positions"] = circulant_positions(n)
        sd[pre + "mlp.0.weight"], sd[pre + "mlp.0.bias"] = xavier(mlp, dim), torch.zeros(mlp)
        sd[pre + "mlp.3.weight"], sd[pre + "mlp.3.bias"] = xavier(dim, mlp), torch.zeros(dim)
        for ln in ("norm1", "norm2"):
            sd[pre + ln + ".weight"], sd[pre + ln + ".bias"] = torch.ones(dim), torch.zeros(dim)
    sd["mlp_head.0.weight"], sd["mlp_head.0.bias"] = torch.ones(dim), torch.zeros(dim)
    sd["mlp_head.1.weight"] = xavier(cfg["num_classes"], dim)
    sd["mlp_head.1.bias"] = torch.zeros(cfg["num_classes"])
    return {k: v.to(dtype) if v.is_floating_point() else v for k, v in sd.items()}


def trainable_keys(sd: Dict[str, torch.Tensor]):
    return [k for k in sd if not k.endswith(_BUFFER_SUFFIXES)]


class CpuTrainer:
    """The reference's training step on CPU: forward (train mode, dropout as configured), CrossEntropy,
    backward, Adam(lr=1e-3) (experiments/utils/training.py:53-69,304-309).  Used only as the reported CPU
    baseline."""

    def __init__(self, model_name: str, cfg: dict, seed: int = 0, lr: float = 1e-3):
        self.model_name, self.cfg = model_name, cfg
        self.dropout = float(cfg.get("dropout", 0.0))
        self.sd = init_state(model_name, cfg, seed)
        self.params = [self.sd[k].requires_grad_(True) for k in trainable_keys(self.sd)]
        self.opt = torch.optim.Adam(self.params, lr=lr)

    def step(self, images: torch.Tensor, labels: torch.Tensor) -> float:
        logits = vit_forward(self.sd, images, self.model_name, self.cfg, dropout=self.dropout)
        loss = F.cross_entropy(logits, labels)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.item())
