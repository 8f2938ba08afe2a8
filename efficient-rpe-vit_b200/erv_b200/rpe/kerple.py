"""KERPLE / "most general" RPE: learnable Toeplitz bias for kernelized attention (reference: models/rpe/kerple.py).

The attention plugins consume `rel_pos_bias` directly inside the Toeplitz-masked tile kernel; `apply_rpe_fft`
keeps the reference's public helper (it materialises D1/D2 and is not used on the model path)."""
from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from .base import BaseRPE


class KERPLEPositionalEncoding(BaseRPE):
    def __init__(self, num_patches: int, dim: int, heads: int):
        super().__init__(num_patches, dim, heads)
        self.max_rel_pos = 2 * num_patches - 1
        # index j - i + (n-1); kerple.py:62-73
        self.rel_pos_bias = nn.Parameter(torch.zeros(heads, self.max_rel_pos))
        nn.init.normal_(self.rel_pos_bias, mean=0.0, std=0.02)

    def _check(self, heads: int, n: int):
        assert heads == self.heads, f"Expected {self.heads} heads, got {heads}"  # kerple.py:155
        assert n == self.num_patches, \
            f"Matrix height {n} doesn't match expected {self.num_patches} from coefficients {self.max_rel_pos}"

    def forward(self, x: torch.Tensor, attention_scores: Optional[torch.Tensor] = None) -> torch.Tensor:
        raise NotImplementedError(  # kerple.py:77-97
            "KERPLE does not use the standard forward() interface. Use apply_rpe_fft() method instead, "
            "which must be called from within kernelized attention computation (FAVOR+/ReLU).")

    def apply_rpe_fft(self, k_prime: torch.Tensor, v: Optional[torch.Tensor] = None) -> torch.Tensor:
        """D1[i] = sum_j c[j-i] phi(k_j)^T v_j ([B,H,N,M,Dh]) with v, else D2[i] = sum_j c[j-i] phi(k_j) ([B,H,N,M]);
        c = exp(rel_pos_bias) (kerple.py:99-344)."""
        b, h, n, m = k_prime.shape
        self._check(h, n)
        c = torch.exp(self.rel_pos_bias)  # [H, 2n-1]; batch p = b*H + h uses row p % H
        if v is None:
            return ops.toeplitz_matmul(c, k_prime.reshape(b * h, n, m)).reshape(b, h, n, m)
        d = v.shape[-1]
        a1 = (k_prime.unsqueeze(-1) * v.unsqueeze(-2)).reshape(b * h, n, m * d)
        return ops.toeplitz_matmul(c, a1).reshape(b, h, n, m, d)

    def extra_repr(self) -> str:
        return (f"num_patches={self.num_patches}, dim={self.dim}, heads={self.heads}, "
                f"max_rel_pos={self.max_rel_pos}, type=KERPLE (Toeplitz tile kernel)")
