"""Common contract of the attention plugins (mirrors models/attention/base.py:14-70 of the reference)."""
from abc import ABC, abstractmethod
from typing import Optional

import torch
import torch.nn as nn


class BaseAttention(ABC, nn.Module):
    """forward(x [B,N,C], mask=None, rpe=None, return_attention=False) -> [B,N,C]."""

    def __init__(self, dim: int, heads: int, dropout: float = 0.0):
        super().__init__()
        assert dim % heads == 0, f"Model dimension {dim} must be divisible by heads {heads}"
        self.dim = dim
        self.heads = heads
        self.head_dim = dim // heads
        self.dropout = dropout
        self.scale = self.head_dim ** -0.5

    @abstractmethod
    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, rpe: Optional[nn.Module] = None,
                return_attention: bool = False) -> torch.Tensor:
        ...

    def extra_repr(self) -> str:
        return f"dim={self.dim}, heads={self.heads}, head_dim={self.head_dim}"
