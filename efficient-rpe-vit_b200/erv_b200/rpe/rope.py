"""1-D RoPE over the raster token index, CLS at position 0 (reference: models/rpe/rope.py)."""
from typing import Optional, Tuple

import torch

from .. import ops
from .base import BaseRPE


class RoPE(BaseRPE):
    def __init__(self, num_patches: int, dim: int, heads: int, theta: float = 10000.0, **kwargs):
        super().__init__(num_patches, dim, heads)
        self.theta = theta
        self.additional_params = kwargs
        # Built with the reference's exact fp32 op sequence (rope.py:53-68) so the caches are bit-identical;
        # they are construction-time constants, non-persistent as in the reference.
        freqs = 1.0 / (theta ** (torch.arange(0, self.head_dim, 2).float() / self.head_dim))
        self.register_buffer("freqs", freqs, persistent=False)
        angles = torch.arange(num_patches).float().unsqueeze(-1) * self.freqs.unsqueeze(0)
        self.register_buffer("cos_cached", torch.cos(angles), persistent=False)
        self.register_buffer("sin_cached", torch.sin(angles), persistent=False)

    def _check(self, heads: int, n: int, head_dim: int):  # rope.py:91-93
        assert head_dim == self.head_dim, f"Expected head_dim={self.head_dim}, got {head_dim}"
        assert heads == self.heads, f"Expected heads={self.heads}, got {heads}"
        assert n <= self.num_patches, f"Sequence length {n} exceeds max {self.num_patches}"

    def apply_rotary_emb(self, q: torch.Tensor, k: torch.Tensor,
                         positions: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Interleaved-pair rotation of q, k [B, H, N, Dh] (rope.py:70-137)."""
        _, h, n, d = q.shape
        self._check(h, n, d)
        if positions is None:
            cos, sin = self.cos_cached[:n], self.sin_cached[:n]
        else:
            idx = positions.to(q.device)
            assert idx.max() < self.num_patches, f"Position {idx.max()} exceeds max {self.num_patches}"
            cos, sin = self.cos_cached[idx], self.sin_cached[idx]
        return ops.rotate(q, ops.ROT_ROPE, cos, sin), ops.rotate(k, ops.ROT_ROPE, cos, sin)

    def forward(self, x: torch.Tensor, attention_scores: Optional[torch.Tensor] = None) -> torch.Tensor:
        return x  # rope.py:139-162: interface filler

    def extra_repr(self) -> str:
        return super().extra_repr() + f", theta={self.theta}"
