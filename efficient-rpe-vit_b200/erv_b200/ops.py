"""torch.autograd bindings of the C-ABI kernels.  Tensors stay torch-owned: the
functions below only hand data pointers, shapes and the current stream to liberv_b200.so."""
import itertools

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _capi as C
from ._capi import FEAT_FAVOR, FEAT_RELU, ROT_CIRCULANT, ROT_NONE, ROT_ROPE  # noqa: F401

_seed_counter = itertools.count(1)

# bench.py sets this to a dict to time the attention C calls with CUDA events on the launching stream
# ({name: [(start, end), ...]}); None (default) adds no work.
PROFILE = None


class _timed:
    def __init__(self, key):
        self.key = key

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.setdefault(self.key, []).append((self.e0, e1))
        return False


def next_seed() -> int:
    """Host-side dropout seed: deterministic under torch.manual_seed, no device sync."""
    return (torch.initial_seed() * 0x9E3779B1 + next(_seed_counter) * 0x85EBCA77) & 0xFFFFFFFFFFFFFFFF


def _f32c(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def _split_dims(qkv: torch.Tensor, heads: int):
    if qkv.dim() != 3 or qkv.shape[-1] % (3 * heads) != 0:
        raise ValueError(f"qkv must be [B, N, 3*heads*head_dim], got {tuple(qkv.shape)} with heads={heads}")
    b, n, c3 = qkv.shape
    return b, n, c3 // 3 // heads


# ---- tables -------------------------------------------------------------------------------------------
class _CirculantTable(torch.autograd.Function):
    """coeffs [H, K, Dh], positions [N-1, K] -> g [H, N, Dh] (row 0 = identity for the CLS token)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, coeffs, positions):
        C.require_cuda(coeffs, positions)
        coeffs = coeffs.to(torch.float32)
        h, k, dh = coeffs.shape
        n = positions.shape[0] + 1
        coeffs, positions = coeffs.contiguous(), positions.to(coeffs.device, torch.float32).contiguous()
        g = torch.empty(h, n, dh, device=coeffs.device, dtype=torch.float32)
        C.check(C.load().erv_circulant_table_fwd(C.ptr(coeffs), C.ptr(positions), h, n, dh, k, C.ptr(g), C.stream()),
                "circulant_table")
        ctx.save_for_backward(coeffs, positions)
        return g

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dg):
        coeffs, positions = ctx.saved_tensors
        h, k, dh = coeffs.shape
        n = positions.shape[0] + 1
        lib = C.load()
        dg = dg.to(torch.float32).contiguous().clone()  # the kernel folds slots in place
        dc = torch.empty_like(coeffs)
        nbytes = lib.erv_circulant_table_bwd_scratch(h, n, dh, k)
        scratch = C.workspace(nbytes, coeffs.device)
        C.check(lib.erv_circulant_table_bwd(C.ptr(coeffs), C.ptr(positions), C.ptr(dg), 1, h, n, dh, k, C.ptr(dc),
                                            C.ptr(scratch), nbytes, C.stream()), "circulant_table_bwd")
        return dc, None


def circulant_table(coeffs, positions):
    return _CirculantTable.apply(coeffs, positions)


def rope_table(theta: float, num_patches: int, head_dim: int, device):
    cos = torch.empty(num_patches, head_dim // 2, device=device, dtype=torch.float32)
    sin = torch.empty_like(cos)
    C.require_cuda(cos)
    C.check(C.load().erv_rope_table(float(theta), num_patches, head_dim, C.ptr(cos), C.ptr(sin), C.stream()), "rope_table")
    return cos, sin


class _Rotate(torch.autograd.Function):
    """Stand-alone rotation of x [B,H,N,Dh]; tab_a/tab_b as in erv_rotate()."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, rot, tab_a, tab_b):
        C.require_cuda(x, tab_a, tab_b)
        x = x.to(torch.float32).contiguous()
        b, h, n, dh = x.shape
        y = torch.empty_like(x)
        tab_a = tab_a.to(torch.float32).contiguous()
        tab_b = None if tab_b is None else tab_b.to(torch.float32).contiguous()
        C.check(C.load().erv_rotate(C.ptr(x), C.ptr(y), b, h, n, dh, rot, C.ptr(tab_a), C.ptr(tab_b), 0, C.stream()), "rotate")
        ctx.rot = rot
        ctx.save_for_backward(x, tab_a, tab_b)
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, tab_a, tab_b = ctx.saved_tensors
        b, h, n, dh = x.shape
        dy = dy.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        lib = C.load()
        C.check(lib.erv_rotate(C.ptr(dy), C.ptr(dx), b, h, n, dh, ctx.rot, C.ptr(tab_a), C.ptr(tab_b), 1, C.stream()), "rotate_bwd")
        dtab = None
        if ctx.rot == ROT_CIRCULANT and ctx.needs_input_grad[2]:
            dtab = torch.zeros_like(tab_a)
            C.check(lib.erv_rotate_table_grad(C.ptr(x), C.ptr(dy), b, h, n, dh, C.ptr(dtab), C.stream()), "rotate_table_grad")
        return dx, None, dtab, None


def rotate(x, rot, tab_a, tab_b=None):
    out = _Rotate.apply(x, rot, tab_a, tab_b)
    return out.to(x.dtype)


# ---- feature maps --------------------------------------------------------------------------------------
class _FeatureMap(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, omega, kind):
        C.require_cuda(x, omega)
        x, omega = x.to(torch.float32).contiguous(), omega.to(torch.float32).contiguous()
        b, h, n, dh = x.shape
        m = omega.shape[-1]
        lib = C.load()
        phi = torch.empty(b, h, n, m, device=x.device, dtype=torch.float32)
        nbytes = lib.erv_feature_map_workspace(h, dh, m)
        ws = C.workspace(nbytes, x.device)
        C.check(lib.erv_feature_map_fwd(C.ptr(x), C.ptr(omega), b, h, n, dh, m, kind, C.ptr(phi), C.ptr(ws), nbytes,
                                        C.stream()), "feature_map")
        ctx.kind = kind
        ctx.save_for_backward(x, omega)
        return phi

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dphi):
        x, omega = ctx.saved_tensors
        b, h, n, dh = x.shape
        m = omega.shape[-1]
        lib = C.load()
        dphi = dphi.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        nbytes = lib.erv_feature_map_workspace(h, dh, m)
        ws = C.workspace(nbytes, x.device)
        C.check(lib.erv_feature_map_bwd(C.ptr(x), C.ptr(omega), C.ptr(dphi), b, h, n, dh, m, ctx.kind, C.ptr(dx),
                                        C.ptr(ws), nbytes, C.stream()), "feature_map_bwd")
        return dx, None, None


def feature_map(x, omega, kind):
    return _FeatureMap.apply(x, omega, kind)


SAVE_KV_STATE = True  # False: the backward rebuilds [S|z] itself (no extra saved tensor)


# ---- attention cores -----------------------------------------------------------------------------------
class _LinearAttention(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, qkv, omega, gtab, heads, kind, rot, tab_a, tab_b):
        # rot == CIRCULANT: gtab (differentiable) is the table; otherwise tab_a/tab_b (buffers)
        C.require_cuda(qkv, omega)
        qkv = qkv.contiguous()
        b, n, dh = _split_dims(qkv, heads)
        m = omega.shape[-1]
        omega = _f32c(omega)
        ta = _f32c(gtab) if rot == ROT_CIRCULANT else _f32c(tab_a)
        tb = _f32c(tab_b)
        lib = C.load()
        out = torch.empty(b, n, heads * dh, device=qkv.device, dtype=qkv.dtype)
        nbytes = lib.erv_linear_attention_workspace(b, n, heads, dh, m, rot, 0)
        ws = C.workspace(nbytes, qkv.device)
        # [phi(k)^T v | sum phi(k)] per (batch, head): written by the forward when a backward will follow
        nstate = lib.erv_linear_attention_state_floats(b, n, heads, dh, m) if (ctx.needs_input_grad[0] and SAVE_KV_STATE) else 0
        state = torch.empty(nstate, device=qkv.device, dtype=torch.float32) if nstate else None
        with _timed("linear_attention_fwd"):
            C.check(lib.erv_linear_attention_fwd(C.ptr(qkv), C.ptr(out), C.ptr(omega), b, n, heads, dh, m, kind, rot,
                                                 C.ptr(ta), C.ptr(tb), C.dtype_code(qkv), C.ptr(state), C.ptr(ws), nbytes,
                                                 C.stream()),
                    "linear_attention")
        ctx.meta = (b, n, heads, dh, m, kind, rot)
        ctx.save_for_backward(qkv, out, omega, ta, tb, state)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        qkv, out, omega, ta, tb, state = ctx.saved_tensors
        b, n, heads, dh, m, kind, rot = ctx.meta
        lib = C.load()
        dout = dout.to(qkv.dtype).contiguous()
        dqkv = torch.empty_like(qkv)
        dg_part = None
        if rot == ROT_CIRCULANT:
            slots = lib.erv_circulant_slots(b, heads)
            dg_part = torch.empty(heads, slots, n, dh, device=qkv.device, dtype=torch.float32)
        nbytes = lib.erv_linear_attention_workspace(b, n, heads, dh, m, rot, 1)
        ws = C.workspace(nbytes, qkv.device)
        with _timed("linear_attention_bwd"):
            C.check(lib.erv_linear_attention_bwd(C.ptr(qkv), C.ptr(out), C.ptr(dout), C.ptr(dqkv), C.ptr(omega), b, n, heads,
                                                 dh, m, kind, rot, C.ptr(ta), C.ptr(tb), C.ptr(dg_part), C.dtype_code(qkv),
                                                 C.ptr(state), C.ptr(ws), nbytes, C.stream()), "linear_attention_bwd")
        dgtab = dg_part.sum(dim=1) if (dg_part is not None and ctx.needs_input_grad[2]) else None
        return dqkv, None, dgtab, None, None, None, None, None


def linear_attention(qkv, omega, heads, kind, rot=ROT_NONE, gtab=None, tab_a=None, tab_b=None):
    return _LinearAttention.apply(qkv, omega, gtab, heads, kind, rot, tab_a, tab_b)


class _KerpleAttention(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, qkv, omega, bias, heads, kind):
        C.require_cuda(qkv, omega, bias)
        qkv = qkv.contiguous()
        b, n, dh = _split_dims(qkv, heads)
        m = omega.shape[-1]
        omega, bias = _f32c(omega), _f32c(bias)
        lib = C.load()
        out = torch.empty(b, n, heads * dh, device=qkv.device, dtype=qkv.dtype)
        den = torch.empty(b, heads, n, device=qkv.device, dtype=torch.float32)
        nbytes = lib.erv_kerple_attention_workspace(b, n, heads, dh, m, 0)
        ws = C.workspace(nbytes, qkv.device)
        with _timed("kerple_attention_fwd"):
            C.check(lib.erv_kerple_attention_fwd(C.ptr(qkv), C.ptr(out), C.ptr(den), C.ptr(omega), C.ptr(bias), b, n, heads,
                                                 dh, m, kind, C.dtype_code(qkv), C.ptr(ws), nbytes, C.stream()),
                    "kerple_attention")
        ctx.meta = (b, n, heads, dh, m, kind)
        ctx.save_for_backward(qkv, out, den, omega, bias)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        qkv, out, den, omega, bias = ctx.saved_tensors
        b, n, heads, dh, m, kind = ctx.meta
        lib = C.load()
        dout = dout.to(qkv.dtype).contiguous()
        dqkv = torch.empty_like(qkv)
        dbias = torch.empty_like(bias)
        nbytes = lib.erv_kerple_attention_workspace(b, n, heads, dh, m, 1)
        ws = C.workspace(nbytes, qkv.device)
        with _timed("kerple_attention_bwd"):
            C.check(lib.erv_kerple_attention_bwd(C.ptr(qkv), C.ptr(out), C.ptr(den), C.ptr(dout), C.ptr(dqkv), C.ptr(dbias),
                                                 C.ptr(omega), C.ptr(bias), b, n, heads, dh, m, kind, C.dtype_code(qkv),
                                                 C.ptr(ws), nbytes, C.stream()), "kerple_attention_bwd")
        return dqkv, None, dbias, None, None


def kerple_attention(qkv, omega, bias, heads, kind):
    return _KerpleAttention.apply(qkv, omega, bias, heads, kind)


class _SoftmaxAttention(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, qkv, gtab, heads, rot, tab_a, tab_b, mask, dropout_p, seed, want_attn):
        C.require_cuda(qkv, mask)
        qkv = qkv.contiguous()
        b, n, dh = _split_dims(qkv, heads)
        ta = _f32c(gtab) if rot == ROT_CIRCULANT else _f32c(tab_a)
        tb = _f32c(tab_b)
        mask8 = None
        if mask is not None:
            if mask.dim() == 4:
                if mask.shape[1] != 1:
                    raise ValueError("mask must be [B, N, N] or [B, 1, N, N]")
                mask = mask[:, 0]
            mask8 = (mask != 0).expand(b, n, n).to(torch.uint8).contiguous()
        lib = C.load()
        out = torch.empty(b, n, heads * dh, device=qkv.device, dtype=qkv.dtype)
        lse = torch.empty(b, heads, n, device=qkv.device, dtype=torch.float32)
        attn = torch.empty(b, heads, n, n, device=qkv.device, dtype=torch.float32) if want_attn else None
        nbytes = lib.erv_softmax_attention_workspace(b, n, heads, dh, rot, 0)
        ws = C.workspace(nbytes, qkv.device)
        seed_dev = seed if torch.is_tensor(seed) else None  # device-resident seed (ops.dropout_seed): graph-replay safe
        seed = 0 if seed_dev is not None else int(seed)
        with _timed("softmax_attention_fwd"):
            C.check(lib.erv_softmax_attention_fwd(C.ptr(qkv), C.ptr(out), C.ptr(lse), C.ptr(attn), C.ptr(mask8), b, n, heads,
                                                  dh, rot, C.ptr(ta), C.ptr(tb), float(dropout_p), seed, C.ptr(seed_dev),
                                                  C.dtype_code(qkv), C.ptr(ws), nbytes, C.stream()), "softmax_attention")
        ctx.meta = (b, n, heads, dh, rot, float(dropout_p), seed)
        ctx.save_for_backward(qkv, out, lse, ta, tb, mask8, seed_dev)
        if want_attn:
            ctx.mark_non_differentiable(attn)
            return out, attn
        return out, None

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout, _dattn):
        qkv, out, lse, ta, tb, mask8, seed_dev = ctx.saved_tensors
        b, n, heads, dh, rot, p, seed = ctx.meta
        lib = C.load()
        dout = dout.to(qkv.dtype).contiguous()
        dqkv = torch.empty_like(qkv)
        dg_part = None
        if rot == ROT_CIRCULANT:
            slots = lib.erv_circulant_slots(b, heads)
            dg_part = torch.empty(heads, slots, n, dh, device=qkv.device, dtype=torch.float32)
        nbytes = lib.erv_softmax_attention_workspace(b, n, heads, dh, rot, 1)
        ws = C.workspace(nbytes, qkv.device)
        with _timed("softmax_attention_bwd"):
            C.check(lib.erv_softmax_attention_bwd(C.ptr(qkv), C.ptr(out), C.ptr(lse), C.ptr(dout), C.ptr(dqkv), C.ptr(mask8),
                                                  b, n, heads, dh, rot, C.ptr(ta), C.ptr(tb), C.ptr(dg_part), p, seed,
                                                  C.ptr(seed_dev), C.dtype_code(qkv), C.ptr(ws), nbytes, C.stream()),
                    "softmax_attention_bwd")
        dgtab = dg_part.sum(dim=1) if (dg_part is not None and ctx.needs_input_grad[1]) else None
        return dqkv, dgtab, None, None, None, None, None, None, None, None


def softmax_attention(qkv, heads, rot=ROT_NONE, gtab=None, tab_a=None, tab_b=None, mask=None, dropout_p=0.0,
                      seed=0, want_attn=False):
    return _SoftmaxAttention.apply(qkv, gtab, heads, rot, tab_a, tab_b, mask, dropout_p, seed, want_attn)


# ---- Toeplitz product ----------------------------------------------------------------------------------
class _Toeplitz(torch.autograd.Function):
    """c [R, 2n-1], x [P, n, d]; batch p uses row p (R == P) or p % R."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, c, x):
        C.require_cuda(c, x)
        c, x = c.to(torch.float32).contiguous(), x.to(torch.float32).contiguous()
        p, n, d = x.shape
        y = torch.empty_like(x)
        C.check(C.load().erv_toeplitz_matmul_fwd(C.ptr(c), C.ptr(x), C.ptr(y), p, c.shape[0], n, d, C.stream()), "toeplitz")
        ctx.save_for_backward(c, x)
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        c, x = ctx.saved_tensors
        p, n, d = x.shape
        dy = dy.to(torch.float32).contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[1] else None
        dc = torch.empty_like(c) if ctx.needs_input_grad[0] else None
        C.check(C.load().erv_toeplitz_matmul_bwd(C.ptr(c), C.ptr(x), C.ptr(dy), C.ptr(dx), C.ptr(dc), p, c.shape[0], n, d,
                                                 C.stream()), "toeplitz_bwd")
        return dc, dx


def toeplitz_matmul(c, x):
    return _Toeplitz.apply(c, x)


# ---- the rest of the block: Linear weight/bias gradient, LayerNorm -------------------------------
class _Linear(torch.autograd.Function):
    """y = x W^T + b.  Forward and dx stay library GEMMs; dW/db use the token-split kernel when the shape fits it."""

    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        o, i = weight.shape
        dy2 = dy.reshape(-1, o)
        x2 = x.reshape(-1, i)
        if x2.dtype != dy2.dtype:
            x2 = x2.to(dy2.dtype)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = (dy2 @ weight.to(dy2.dtype)).reshape(x.shape).to(x.dtype)
        r = dy2.shape[0]
        lib = C.load() if dy2.is_cuda else None
        if (lib is not None and dy2.dtype in (torch.float32, torch.bfloat16) and weight.dtype == torch.float32
                and lib.erv_linear_wgrad_supported(r, o, i)):
            dy2, x2 = dy2.contiguous(), x2.contiguous()
            dw = torch.empty_like(weight)
            db = torch.empty(o, device=weight.device, dtype=torch.float32) if ctx.has_bias else None
            nbytes = lib.erv_linear_wgrad_workspace(r, o, i)
            ws = C.workspace(nbytes, dy2.device)
            C.check(lib.erv_linear_wgrad(C.ptr(dy2), C.ptr(x2), C.ptr(dw), C.ptr(db), r, o, i, C.dtype_code(dy2), C.ptr(ws),
                                         nbytes, C.stream()), "linear_wgrad")
        else:
            if ctx.needs_input_grad[1]:
                dw = (dy2.t() @ x2).to(weight.dtype)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = dy2.sum(dim=0).to(weight.dtype)
        return dx, dw, db


def linear(x, weight, bias=None):
    return _Linear.apply(x, weight, bias)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias, eps):
        shape = x.shape
        c = shape[-1]
        x2 = x.reshape(-1, c).contiguous()
        r = x2.shape[0]
        y = torch.empty_like(x2)
        mean = torch.empty(r, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        C.check(C.load().erv_layernorm_fwd(C.ptr(x2), C.ptr(weight), C.ptr(bias), C.ptr(y), C.ptr(mean), C.ptr(rstd), r, c,
                                           float(eps), C.stream()), "layernorm")
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.shape = shape
        return y.reshape(shape)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x2, weight, mean, rstd = ctx.saved_tensors
        r, c = x2.shape
        dy2 = dy.reshape(r, c).to(torch.float32).contiguous()
        lib = C.load()
        dx = torch.empty_like(x2)
        dg, db = torch.empty_like(weight), torch.empty_like(weight)
        nbytes = lib.erv_layernorm_bwd_workspace(r, c)
        ws = C.workspace(nbytes, x2.device)
        C.check(lib.erv_layernorm_bwd(C.ptr(dy2), C.ptr(x2), C.ptr(weight), C.ptr(mean), C.ptr(rstd), C.ptr(dx), C.ptr(dg),
                                      C.ptr(db), r, c, C.ptr(ws), nbytes, C.stream()), "layernorm_bwd")
        return dx.reshape(ctx.shape), dg, db, None


# ---- fused block kernels (dim 32, MLP width 64): LayerNorm1 + qkv ; proj + residual + LayerNorm2 + MLP + residual ------
FUSED_BLOCK = True  # False: the block runs op by op (Linear / LayerNorm / GELU / Dropout modules)
# Fused patch embedding + CLS + positions: tcgen05 kernels for 4x4 patches (csrc/erv_block_tc.cu), plain FMA kernels with
# per-token gathers otherwise (csrc/erv_embed_head.cu; slower than the library chain they replace, so only used on request).
FUSED_EMBED = True


def embed_fast(cin: int, patch: int, n_tokens: int) -> bool:
    """True when the tcgen05 embedding kernels cover this geometry."""
    return patch == 4 and 1 <= cin <= 4 and n_tokens <= 128
# True (set by erv_b200.train.Trainer): the fused backward kernels add parameter gradients straight into the existing
# fp32 `.grad` buffers and return None to autograd, which saves one accumulation kernel per parameter.  Only valid when
# nothing else looks at the per-call gradients (no hooks, no create_graph); off by default.
GRAD_INPLACE = False


def _grad_targets(params):
    """.grad buffers of the saved parameters when fused accumulation applies, else None."""
    if not GRAD_INPLACE:
        return None
    out = []
    for t in params:
        if t is None:
            out.append(None)
            continue
        g = t.grad
        if g is None or g.dtype != torch.float32 or not g.is_contiguous() or not t.requires_grad:
            return None
        out.append(g)
    return out


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])

_drop_state = {}


def dropout_seed(device):
    """A fresh device-resident dropout seed (int64 scalar): a per-device counter that lives on the GPU, so the value
    changes on every CUDA-graph replay.  Initialised from torch's seed and the data-parallel rank."""
    st = _drop_state.get(device)
    if st is None:
        rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
        st = torch.tensor([(torch.initial_seed() * 0x9E3779B1 + rank * 0x632BE5AB) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64,
                          device=device)
        _drop_state[device] = st
    seed = st.clone()
    st.add_(0x2545F491)
    return seed


class _BlockLnQkv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w, b, eps, out_dtype):
        C.require_cuda(x, ln_w, ln_b, w)
        dim = x.shape[-1]
        x2 = x.reshape(-1, dim).contiguous()
        rows = x2.shape[0]
        lib = C.load()
        # bf16 autocast: the kernel writes the bf16 qkv an autocast Linear would hand the attention core (no cast kernel)
        direct = out_dtype == torch.bfloat16 and lib.erv_block_act_bf16_supported()
        qkv = torch.empty(rows, 3 * dim, device=x.device, dtype=torch.bfloat16 if direct else torch.float32)
        C.check(lib.erv_block_ln_qkv_fwd(C.ptr(x2), C.ptr(ln_w), C.ptr(ln_b), C.ptr(w), C.ptr(b), C.ptr(qkv), C.dtype_code(qkv),
                                         rows, dim, float(eps), C.stream()), "block_ln_qkv")
        if out_dtype is not None and qkv.dtype != out_dtype:
            qkv = qkv.to(out_dtype)
        ctx.save_for_backward(x2, ln_w, ln_b, w, b)
        ctx.meta = (x.shape, float(eps), b is not None)
        # the block input is handed on as a second output: the residual branch reads it from here, so its gradient arrives in
        # this backward and is added inside the kernel (dres) instead of by an autograd accumulation kernel
        return qkv.reshape(*x.shape[:-1], 3 * dim), x.view_as(x)

    @staticmethod
    def backward(ctx, dqkv, dres):
        x2, ln_w, ln_b, w, b = ctx.saved_tensors
        shape, eps, has_bias = ctx.meta
        rows, dim = x2.shape
        lib = C.load()
        if dqkv is None:
            dqkv = torch.zeros(rows, 3 * dim, device=x2.device, dtype=torch.float32)
        dq = dqkv.reshape(rows, 3 * dim)
        if not (dq.dtype == torch.bfloat16 and lib.erv_block_act_bf16_supported()):
            dq = dq.to(torch.float32)
        dq = dq.contiguous()
        if dres is not None:
            dres = dres.reshape(rows, dim).to(torch.float32).contiguous()
        dx = torch.empty_like(x2)
        nbytes = lib.erv_block_ln_qkv_bwd_workspace(rows)
        ws = C.workspace(nbytes, x2.device)
        tgt = _grad_targets((w, b, ln_w, ln_b))
        dpar = None if tgt else torch.empty(lib.erv_block_ln_qkv_params(), device=x2.device, dtype=torch.float32)
        C.check(lib.erv_block_ln_qkv_bwd(C.ptr(x2), C.ptr(dq), C.dtype_code(dq), C.ptr(dres), C.ptr(ln_w), C.ptr(ln_b), C.ptr(w), C.ptr(dx), C.ptr(dpar),
                                         _ptr_array(tgt) if tgt else None, rows, dim, eps, C.ptr(ws), nbytes, C.stream()),
                "block_ln_qkv_bwd")
        if tgt:
            return dx.reshape(shape), None, None, None, None, None, None
        nw = 3 * dim * dim
        dw = dpar[:nw].view(3 * dim, dim)
        db = dpar[nw:nw + 3 * dim] if has_bias else None
        return dx.reshape(shape), dpar[nw + 3 * dim:nw + 4 * dim], dpar[nw + 4 * dim:nw + 5 * dim], dw, db, None, None


def block_supported(dim: int, mlp_dim: int) -> bool:
    return bool(C.load().erv_block_supported(int(dim), int(mlp_dim)))


def block_ln_qkv(x, ln_w, ln_b, w, b, eps=1e-5, with_residual=False, out_dtype=None):
    """qkv = LayerNorm(x) W^T (+ b), fp32 or (out_dtype=torch.bfloat16, autocast) bf16.  with_residual=True also returns x itself
    for the residual branch (see _BlockLnQkv)."""
    qkv, xr = _BlockLnQkv.apply(x, ln_w, ln_b, w, b, eps, out_dtype)
    return (qkv, xr) if with_residual else qkv


def _param_array(params):
    return (C.c_void_p * len(params))(*[t.data_ptr() for t in params])


class _BlockMlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, x, wp, bp, ln_w, ln_b, w1, b1, w2, b2, eps, p_drop, seed, salt):
        C.require_cuda(a, x, wp)
        dim, mlp_dim = x.shape[-1], w1.shape[0]
        lib = C.load()
        a2 = a.reshape(-1, dim)
        if not (a2.dtype == torch.bfloat16 and lib.erv_block_act_bf16_supported()):
            a2 = a2.to(torch.float32)
        a2 = a2.contiguous()
        x2 = x.reshape(-1, dim).contiguous()
        rows = x2.shape[0]
        params = (wp, bp, ln_w, ln_b, w1, b1, w2, b2)
        y = torch.empty_like(x2)
        C.check(lib.erv_block_mlp_fwd(C.ptr(a2), C.dtype_code(a2), C.ptr(x2), _param_array(params), C.ptr(y), rows, dim, mlp_dim,
                                      float(eps), float(p_drop), C.ptr(seed), int(salt), C.stream()), "block_mlp")
        ctx.save_for_backward(a2, x2, seed, *params)
        ctx.meta = (x.shape, float(eps), float(p_drop), int(salt), mlp_dim)
        ctx.a_dtype = a.dtype
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        a2, x2, seed, *params = ctx.saved_tensors
        shape, eps, p_drop, salt, mlp_dim = ctx.meta
        rows, dim = x2.shape
        lib = C.load()
        dy2 = dy.reshape(rows, dim).to(torch.float32).contiguous()
        da, dx1 = torch.empty_like(a2), torch.empty_like(x2)
        nbytes = lib.erv_block_mlp_bwd_workspace(rows)
        ws = C.workspace(nbytes, x2.device)
        tgt = _grad_targets(params)
        dpar = None if tgt else torch.empty(lib.erv_block_mlp_params(), device=x2.device, dtype=torch.float32)
        C.check(lib.erv_block_mlp_bwd(C.ptr(a2), C.dtype_code(a2), C.ptr(x2), C.ptr(dy2), _param_array(params), C.ptr(da), C.ptr(dx1), C.ptr(dpar),
                                      _ptr_array(tgt) if tgt else None, rows, dim, mlp_dim, eps, p_drop, C.ptr(seed), salt,
                                      C.ptr(ws), nbytes, C.stream()), "block_mlp_bwd")
        da = da.reshape(shape).to(ctx.a_dtype)
        if tgt:
            return (da, dx1.reshape(shape), *([None] * 12))
        o, grads = 0, []
        for t in params:  # dparams follows the parameter order
            grads.append(dpar[o:o + t.numel()].view(t.shape))
            o += t.numel()
        return (da, dx1.reshape(shape), *grads, None, None, None, None)


def block_mlp(a, x, wp, bp, ln_w, ln_b, w1, b1, w2, b2, eps=1e-5, p_drop=0.0, seed=None, salt=0):
    return _BlockMlp.apply(a, x, wp, bp, ln_w, ln_b, w1, b1, w2, b2, eps, p_drop, seed, salt)


# ---- the two ends of the ViT: patch embedding + CLS + positions ; final LayerNorm + classifier + cross-entropy -------------
class _Embed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, w, b, cls, pos, patch):
        C.require_cuda(images, w)
        images = images.contiguous()
        bsz, cin, s, _ = images.shape
        n = (s // patch) ** 2 + 1
        out = torch.empty(bsz, n, w.shape[0], device=images.device, dtype=torch.float32)
        C.check(C.load().erv_embed_fwd(C.ptr(images), C.ptr(w), C.ptr(b), C.ptr(cls), C.ptr(pos), C.ptr(out), bsz, cin, s, patch,
                                       C.stream()), "embed")
        ctx.save_for_backward(images, w, b, cls, pos)  # cls [1,1,dim], pos [1,N,dim]: the parameters themselves
        ctx.patch = patch
        return out

    @staticmethod
    def backward(ctx, dout):
        images, w, b, cls, pos = ctx.saved_tensors
        bsz, cin, s, _ = images.shape
        lib = C.load()
        dout = dout.to(torch.float32).contiguous()
        tgt = _grad_targets((w, b, cls, pos))
        grads = tgt if tgt else [torch.empty_like(t) for t in (w, b, cls, pos)]
        nbytes = lib.erv_embed_bwd_workspace(bsz, cin, s, ctx.patch)
        ws = C.workspace(nbytes, images.device)
        C.check(lib.erv_embed_bwd(C.ptr(images), C.ptr(dout), *[C.ptr(g) for g in grads], 1 if tgt else 0, bsz, cin, s, ctx.patch,
                                  C.ptr(ws), nbytes, C.stream()), "embed_bwd")
        if tgt:
            return None, None, None, None, None, None
        return (None, *grads, None)


def embed(images, w, b, cls, pos, patch):
    return _Embed.apply(images, w, b, cls, pos, patch)


def embed_supported(dim: int, patch_dim: int) -> bool:
    return bool(C.load().erv_embed_supported(int(dim), int(patch_dim)))


class _HeadLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w, b, labels, eps):
        C.require_cuda(x, w, labels)
        x = x.contiguous()
        bsz, n, dim = x.shape
        labels = labels.to(torch.int64).contiguous()
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        lib = C.load()
        nbytes = lib.erv_head_loss_workspace(bsz, w.shape[0])
        ws = C.workspace(nbytes, x.device)
        C.check(lib.erv_head_loss_fwd(C.ptr(x), C.ptr(ln_w), C.ptr(ln_b), C.ptr(w), C.ptr(b), C.ptr(labels), C.ptr(loss), bsz, n,
                                      dim, w.shape[0], float(eps), C.ptr(ws), nbytes, C.stream()), "head_loss")
        ctx.save_for_backward(x, ln_w, ln_b, w, b, labels)
        ctx.eps = float(eps)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        x, ln_w, ln_b, w, b, labels = ctx.saved_tensors
        bsz, n, dim = x.shape
        k = w.shape[0]
        dloss = dloss.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        tgt = _grad_targets((w, b, ln_w, ln_b))
        dpar = None if tgt else torch.empty(k * dim + k + 2 * dim, device=x.device, dtype=torch.float32)
        lib = C.load()
        nbytes = lib.erv_head_loss_workspace(bsz, k)
        ws = C.workspace(nbytes, x.device)
        C.check(lib.erv_head_loss_bwd(C.ptr(x), C.ptr(ln_w), C.ptr(ln_b), C.ptr(w), C.ptr(b), C.ptr(labels), C.ptr(dloss),
                                      C.ptr(dx), C.ptr(dpar), _ptr_array(tgt) if tgt else None, bsz, n, dim, k, ctx.eps,
                                      C.ptr(ws), nbytes, C.stream()), "head_loss_bwd")
        if tgt:
            return dx, None, None, None, None, None, None
        o = k * dim
        return dx, dpar[o + k:o + k + dim], dpar[o + k + dim:], dpar[:o].view(k, dim), dpar[o:o + k], None, None


def head_loss(x, ln_w, ln_b, w, b, labels, eps=1e-5):
    """mean cross-entropy of Linear(LayerNorm(x[:, 0])) against labels, as one kernel (and one for the backward)."""
    return _HeadLoss.apply(x, ln_w, ln_b, w, b, labels, eps)


def layer_norm(x, weight, bias, eps=1e-5):
    if (not x.is_cuda) or x.shape[-1] > 256 or weight is None or bias is None or weight.dtype != torch.float32:
        return torch.nn.functional.layer_norm(x, (x.shape[-1],), weight, bias, eps)
    return _LayerNorm.apply(x, weight, bias, eps)
