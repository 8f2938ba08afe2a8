"""Common contract of the RPE plugins (mirrors models/rpe/base.py:14-81 of the reference)."""
from abc import ABC, abstractmethod
from typing import Optional

import torch
import torch.nn as nn


class BaseRPE(ABC, nn.Module):
    def __init__(self, num_patches: int, dim: int, heads: int):
        super().__init__()
        self.num_patches = num_patches
        self.dim = dim
        self.heads = heads
        self.head_dim = dim // heads

    @abstractmethod
    def forward(self, x: torch.Tensor, attention_scores: Optional[torch.Tensor] = None) -> torch.Tensor:
        ...

    def get_relative_positions(self, seq_len: int) -> torch.Tensor:
        pos = torch.arange(seq_len)
        return pos.unsqueeze(1) - pos.unsqueeze(0)

    def extra_repr(self) -> str:
        return f"num_patches={self.num_patches}, dim={self.dim}, heads={self.heads}"
