"""Circulant-STRING: learnable 2-D rotation, CLS token untouched (reference: models/rpe/circulant_string.py).

The reference rotates with FFT -> multiply by exp(mu) -> IFFT in complex64.  Here the same orthogonal
map is a Dh-point circular convolution with g = Re IFFT(exp(i theta)); the g table [H, N, Dh] is built by
one tiny kernel from the coefficients and consumed inside the attention kernels.
"""
import math
import warnings
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from .base import BaseRPE


class CirculantStringRPE(BaseRPE):
    def __init__(self, num_patches: int, dim: int, heads: int, coord_dim: int = 2,
                 block_size: Optional[int] = None, **kwargs):
        super().__init__(num_patches, dim, heads)
        self.coord_dim = coord_dim
        self.block_size = block_size
        self.additional_params = kwargs  # image_size / patch_size arrive here from the factory and are unused
        if block_size is not None:  # circulant_string.py:128-144
            if self.head_dim % block_size != 0:
                raise ValueError(f"head_dim ({self.head_dim}) must be divisible by block_size ({block_size})")
            self.num_blocks = self.head_dim // block_size
            warnings.warn(f"block_size={block_size} specified but block-circulant optimization not yet "
                          "implemented. Using full-dimension circulant.", UserWarning)
            self.block_size = None
        self.circulant_coeffs = nn.Parameter(torch.zeros(heads, coord_dim, self.head_dim))
        nn.init.normal_(self.circulant_coeffs, mean=0.0, std=0.01)
        self._setup_positions(num_patches)

    def _setup_positions(self, num_patches: int) -> None:
        """[x, y] integer grid, row-major, for the num_patches-1 patch tokens (circulant_string.py:160-205)."""
        n = num_patches - 1
        if n <= 0:
            self.register_buffer("patch_positions", torch.zeros(0, self.coord_dim))
            self._patches_per_side = 0
            return
        side = int(math.sqrt(n))
        if side ** 2 != n:
            raise ValueError(f"num_patches - 1 = {n} must be a perfect square for 2D position encoding. "
                             f"Got sqrt ≈ {math.sqrt(n):.2f}")
        self._patches_per_side = side
        axis = torch.arange(side, dtype=torch.float32)
        yy, xx = torch.meshgrid(axis, axis, indexing="ij")
        self.register_buffer("patch_positions", torch.stack([xx.flatten(), yy.flatten()], dim=-1))

    def get_eigenvalues(self) -> torch.Tensor:
        """lambda(C - C^T) = FFT(c) - conj FFT(c) = 2i Im FFT(c) (circulant_string.py:207-232); diagnostic only."""
        lam = torch.fft.fft(self.circulant_coeffs, dim=-1)
        return lam - torch.conj(lam)

    def rotation_table(self, n: int, heads: int, head_dim: int) -> torch.Tensor:
        """g [H, n, Dh] for a sequence of n tokens (CLS + n-1 grid patches); differentiable wrt the coefficients."""
        assert heads == self.heads and head_dim == self.head_dim, \
            f"Expected heads={self.heads}, head_dim={self.head_dim}, got {heads}, {head_dim}"
        assert n - 1 == self.patch_positions.shape[0], \
            f"Sequence has {n - 1} patch tokens, position grid has {self.patch_positions.shape[0]}"
        return ops.circulant_table(self.circulant_coeffs, self.patch_positions)

    def apply_rotation(self, x: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        """Rotate patch tokens x [B, H, N, Dh] (no CLS row) sitting at `positions` [N, coord_dim]."""
        g = ops.circulant_table(self.circulant_coeffs, positions.to(x.device))[:, 1:, :]
        return ops.rotate(x, ops.ROT_CIRCULANT, g)

    def apply_circulant_string(self, q: torch.Tensor, k: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Rotate q, k [B, H, N, Dh]; index 0 (CLS) passes through (circulant_string.py:297-341)."""
        if q.shape[2] <= 1:
            return q, k
        g = self.rotation_table(q.shape[2], q.shape[1], q.shape[3])
        return ops.rotate(q, ops.ROT_CIRCULANT, g), ops.rotate(k, ops.ROT_CIRCULANT, g)

    def forward(self, x: torch.Tensor, attention_scores: Optional[torch.Tensor] = None) -> torch.Tensor:
        return x  # circulant_string.py:343-365: interface filler

    def extra_repr(self) -> str:
        return (f"{super().extra_repr()}, coord_dim={self.coord_dim}, patches_per_side={self._patches_per_side}, "
                f"params_per_head={self.coord_dim * self.head_dim}")
