"""Model factory: the 11 usable variants by name (reference: models/factory.py).  Names, aliases, config merging
and error behaviour follow the reference; the classes behind the two registries are the B200 plugins."""
from typing import Any, Callable, Dict, Optional

import torch.nn as nn

from .attention import ATTENTION_REGISTRY
from .rpe import RPE_REGISTRY
from .vit import BaseViT

MODEL_VARIANTS = {
    "baseline": ("softmax", None),
    "baseline_most_general": ("softmax", "most_general"),  # constructible, raises at forward (softmax.py:69-77)
    "baseline_circulant": ("softmax", "circulant_string"),
    "baseline_rope": ("softmax", "rope"),
    "performer_favor": ("favor_plus", None),
    "performer_favor_most_general": ("favor_plus", "most_general"),
    "performer_favor_circulant": ("favor_plus", "circulant_string"),
    "performer_favor_rope": ("favor_plus", "rope"),
    "performer_relu": ("relu", None),
    "performer_relu_most_general": ("relu", "most_general"),
    "performer_relu_circulant": ("relu", "circulant_string"),
    "performer_relu_rope": ("relu", "rope"),
    "performer": ("favor_plus", None),
    "vit": ("softmax", None),
}


def create_attention_builder(attention_type: str, attention_config: Optional[Dict[str, Any]] = None) -> Callable:
    if attention_type not in ATTENTION_REGISTRY:
        raise ValueError(f"Unknown attention type: {attention_type}. Available types: {list(ATTENTION_REGISTRY.keys())}")
    cls, cfg = ATTENTION_REGISTRY[attention_type], attention_config or {}

    def builder(dim: int, heads: int, dropout: float = 0.0) -> nn.Module:
        return cls(dim=dim, heads=heads, dropout=dropout, **cfg)

    return builder


def create_rpe_builder(rpe_type: Optional[str], rpe_config: Optional[Dict[str, Any]] = None,
                       image_size: Optional[int] = None, patch_size: Optional[int] = None) -> Optional[Callable]:
    if rpe_type is None:
        return None
    if rpe_type not in RPE_REGISTRY:
        raise ValueError(f"Unknown RPE type: {rpe_type}. Available types: {list(RPE_REGISTRY.keys())}")
    cls, cfg = RPE_REGISTRY[rpe_type], rpe_config or {}
    if rpe_type in ("circulant_string", "circulant") and image_size is not None and patch_size is not None:
        cfg = dict(cfg, image_size=image_size, patch_size=patch_size)  # factory.py:109-112

    def builder(num_patches: int, dim: int, heads: int) -> nn.Module:
        return cls(num_patches=num_patches, dim=dim, heads=heads, **cfg)

    return builder


def create_model(model_name: str, dataset_config: Dict[str, Any], attention_config: Optional[Dict[str, Any]] = None,
                 rpe_config: Optional[Dict[str, Any]] = None, **kwargs) -> BaseViT:
    """create_model('performer_favor', CIFAR10_CONFIG, attention_config={'num_features': 256}, patch_size=4)."""
    if model_name in MODEL_VARIANTS:
        attention_type, rpe_type = MODEL_VARIANTS[model_name]
    else:  # "<attention>_<rpe>" spelling (factory.py:168-185)
        head, _, tail = model_name.partition("_")
        attention_type, rpe_type = head, (tail or None)
        if attention_type not in ATTENTION_REGISTRY:
            raise ValueError(f"Unknown model: {model_name}. Available models: {list(MODEL_VARIANTS.keys())}")
    config = dict(dataset_config)
    config.update(kwargs)
    if "attention_params" in config:
        merged = dict(config.pop("attention_params").get(attention_type, {}))  # copy: do not mutate the caller's dict
        merged.update(attention_config or {})
        attention_config = merged
    if "rpe_params" in config and rpe_type:
        merged = dict(config.pop("rpe_params").get(rpe_type, {}))
        merged.update(rpe_config or {})
        rpe_config = merged
    model = BaseViT(
        image_size=config["image_size"], in_channels=config["in_channels"], patch_size=config["patch_size"],
        num_classes=config["num_classes"], dim=config["dim"], depth=config["depth"], heads=config["heads"],
        mlp_dim=config["mlp_dim"], dropout=config.get("dropout", 0.1),
        attention_builder=create_attention_builder(attention_type, attention_config),
        rpe_builder=create_rpe_builder(rpe_type, rpe_config, config.get("image_size"), config.get("patch_size")))
    model.model_name, model.attention_type, model.rpe_type = model_name, attention_type, rpe_type
    return model


def list_available_models() -> list:
    return list(MODEL_VARIANTS.keys())


def get_model_info(model_name: str) -> Dict[str, Any]:
    if model_name not in MODEL_VARIANTS:
        raise ValueError(f"Unknown model: {model_name}")
    attention_type, rpe_type = MODEL_VARIANTS[model_name]
    return {"name": model_name, "attention_type": attention_type, "rpe_type": rpe_type,
            "attention_complexity": "O(N²)" if attention_type == "softmax" else "O(N)",
            "has_rpe": rpe_type is not None}
