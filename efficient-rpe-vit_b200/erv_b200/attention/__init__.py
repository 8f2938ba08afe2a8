"""Attention plugins; same registry keys as the reference (models/attention/__init__.py:16-23)."""
from .base import BaseAttention
from .softmax import SoftmaxAttention
from .favor_plus import FAVORPlusAttention
from .relu import ReLUAttention

ATTENTION_REGISTRY = {
    "softmax": SoftmaxAttention,
    "baseline": SoftmaxAttention,
    "favor_plus": FAVORPlusAttention,
    "favor+": FAVORPlusAttention,
    "performer": FAVORPlusAttention,
    "relu": ReLUAttention,
}

__all__ = ["BaseAttention", "SoftmaxAttention", "FAVORPlusAttention", "ReLUAttention", "ATTENTION_REGISTRY"]
