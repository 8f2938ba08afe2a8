"""RPE plugins; same registry keys as the reference (models/rpe/__init__.py:17-24)."""
from .base import BaseRPE
from .kerple import KERPLEPositionalEncoding
from .circulant_string import CirculantStringRPE
from .rope import RoPE

RPE_REGISTRY = {
    "most_general": KERPLEPositionalEncoding,
    "kerple": KERPLEPositionalEncoding,
    "circulant_string": CirculantStringRPE,
    "circulant": CirculantStringRPE,
    "rope": RoPE,
    "rotary": RoPE,
}

__all__ = ["BaseRPE", "KERPLEPositionalEncoding", "CirculantStringRPE", "RoPE", "RPE_REGISTRY"]
