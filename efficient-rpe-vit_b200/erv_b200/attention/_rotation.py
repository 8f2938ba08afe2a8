"""Turns an RPE module into the (rot, gtab, tab_a, tab_b) arguments of the fused kernels."""
from .. import ops
from ..rpe import CirculantStringRPE, RoPE


def rotation_args(rpe, x_shape, heads, head_dim):
    """RoPE -> cached cos/sin rows 0..N-1 (rope.py:95-104); Circulant-STRING -> the g table built from the
    learnable coefficients (circulant_string.py:234-295); anything else -> no rotation."""
    n = x_shape[1]
    if isinstance(rpe, RoPE):
        rpe._check(heads, n, head_dim)
        return ops.ROT_ROPE, None, rpe.cos_cached[:n], rpe.sin_cached[:n]
    if isinstance(rpe, CirculantStringRPE):
        if n <= 1:  # only the CLS token: nothing to rotate (circulant_string.py:318-320)
            return ops.ROT_NONE, None, None, None
        return ops.ROT_CIRCULANT, rpe.rotation_table(n, heads, head_dim), None, None
    return ops.ROT_NONE, None, None, None
