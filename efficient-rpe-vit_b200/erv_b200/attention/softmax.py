"""Softmax attention plugin (reference: models/attention/softmax.py:14-127) on the flash-style tile kernel."""
from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ..nn import Linear
from ..rpe import KERPLEPositionalEncoding
from ._rotation import rotation_args
from .base import BaseAttention


class SoftmaxAttention(BaseAttention):
    def __init__(self, dim: int, heads: int, dropout: float = 0.0, qkv_bias: bool = False):
        super().__init__(dim, heads, dropout)
        self.qkv = Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = Linear(dim, dim)
        self.attn_dropout = nn.Dropout(dropout)  # kept for module-tree parity; the kernel applies it
        self.proj_dropout = nn.Dropout(dropout)

    def before_qkv(self, x_shape, rpe):
        if isinstance(rpe, KERPLEPositionalEncoding):  # softmax.py:69-77
            raise NotImplementedError(
                "KERPLE RPE is designed specifically for kernelized attention (FAVOR+/ReLU Performer) and "
                "cannot be used with standard softmax attention. For softmax attention, use RoPE or "
                "Circulant-STRING RPE instead.")

    def core(self, qkv: torch.Tensor, x_shape, rpe: Optional[nn.Module] = None, mask=None, return_attention: bool = False):
        """Packed qkv [B, N, 3C] -> attention output [B, N, C] before the output projection (and the weights)."""
        rot, gtab, ta, tb = rotation_args(rpe, x_shape, self.heads, self.head_dim)
        p = self.attn_dropout.p if self.training else 0.0
        # the seed lives on the device (a counter advanced by every call), so CUDA-graph replays draw new masks
        out, attn = ops.softmax_attention(qkv, self.heads, rot, gtab, ta, tb, mask, p,
                                          ops.dropout_seed(qkv.device) if p > 0 else 0, return_attention)
        return (out, attn) if return_attention else out

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, rpe: Optional[nn.Module] = None,
                return_attention: bool = False):
        self.before_qkv(x.shape, rpe)
        res = self.core(self.qkv(x), x.shape, rpe, mask, return_attention)
        out, attn = res if return_attention else (res, None)
        out = self.proj_dropout(self.proj(out))
        return (out, attn) if return_attention else out

    def extra_repr(self) -> str:
        return super().extra_repr() + ", complexity=O(N²)"
