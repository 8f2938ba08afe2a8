"""Shared body of the two random-feature attention plugins (FAVOR+ and ReLU).

Reference: models/attention/favor_plus.py:16-282 and relu.py:16-281, which are line-for-line the same
apart from the feature map.  Construction order (qkv, proj, omega, redraw_counter) and RNG use follow
the reference so that the same torch seed gives the same parameters and the same `omega`.
"""
import math
from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ..nn import Linear
from ..rpe import KERPLEPositionalEncoding
from ._rotation import rotation_args
from .base import BaseAttention


class RandomFeatureAttention(BaseAttention):
    _kind = None  # ops.FEAT_*

    def __init__(self, dim: int, heads: int, dropout: float = 0.0, num_features: Optional[int] = None,
                 use_orthogonal: bool = True, feature_redraw_interval: Optional[int] = None,
                 qkv_bias: bool = False):
        super().__init__(dim, heads, dropout)
        if num_features is None:
            num_features = int(self.head_dim * math.log(self.head_dim))  # favor_plus.py:50-53
        self.num_features = num_features
        self.use_orthogonal = use_orthogonal
        self.feature_redraw_interval = feature_redraw_interval
        self.qkv = Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = Linear(dim, dim)
        self.proj_dropout = nn.Dropout(dropout)
        self.register_buffer("omega", self._draw_features())
        self.register_buffer("redraw_counter", torch.tensor(0))

    # favor_plus.py:73-110: iid Gaussian, or QR-orthogonal Dh x Dh blocks scaled by sqrt(Dh)
    def _draw_features(self) -> torch.Tensor:
        h, d, m = self.heads, self.head_dim, self.num_features
        if not self.use_orthogonal:
            return torch.randn(h, d, m)
        per_head = []
        for _ in range(h):
            if m <= d:
                q, _ = torch.linalg.qr(torch.randn(d, m), mode="reduced")
                per_head.append(q * math.sqrt(d))
            else:
                blocks = [torch.linalg.qr(torch.randn(d, d), mode="reduced")[0] for _ in range(math.ceil(m / d))]
                per_head.append(torch.cat(blocks, dim=1)[:, :m] * math.sqrt(d))
        return torch.stack(per_head, dim=0)

    def _create_random_features(self):
        """Redraw in place, on the buffer's device (the reference re-registers a CPU buffer: SURVEY.md appendix C.6)."""
        self.omega.copy_(self._draw_features().to(self.omega.device, self.omega.dtype))

    def _features(self, x: torch.Tensor, omega: torch.Tensor) -> torch.Tensor:
        return ops.feature_map(x, omega, self._kind)

    def before_qkv(self, x_shape, rpe):
        """Per-call bookkeeping that precedes the qkv projection: feature redraw, RPE shape checks."""
        if self.training and self.feature_redraw_interval is not None:  # favor_plus.py:168-171
            if int(self.redraw_counter) % self.feature_redraw_interval == 0:
                self._create_random_features()
            self.redraw_counter += 1
        if isinstance(rpe, KERPLEPositionalEncoding):
            rpe._check(self.heads, x_shape[1])

    def core(self, qkv: torch.Tensor, x_shape, rpe: Optional[nn.Module] = None) -> torch.Tensor:
        """Packed qkv [B, N, 3C] -> attention output [B, N, C] before the output projection."""
        if isinstance(rpe, KERPLEPositionalEncoding):
            return ops.kerple_attention(qkv, self.omega, rpe.rel_pos_bias, self.heads, self._kind)
        rot, gtab, ta, tb = rotation_args(rpe, x_shape, self.heads, self.head_dim)
        return ops.linear_attention(qkv, self.omega, self.heads, self._kind, rot, gtab, ta, tb)

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, rpe: Optional[nn.Module] = None,
                return_attention: bool = False) -> torch.Tensor:
        self.before_qkv(x.shape, rpe)
        out = self.core(self.qkv(x), x.shape, rpe)
        out = self.proj_dropout(self.proj(out))
        if return_attention:  # favor_plus.py:267-273: raised after the work, as in the reference
            raise NotImplementedError(
                f"{type(self).__name__} doesn't compute explicit attention matrices. "
                "Returning attention weights would require O(N²) computation.")
        return out

    def extra_repr(self) -> str:
        return (super().extra_repr() + f", complexity=O(N), num_features={self.num_features}, "
                f"orthogonal={self.use_orthogonal}")
