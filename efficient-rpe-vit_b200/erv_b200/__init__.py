"""erv_b200: B200-native attention hot path of efficient-rpe-vit behind the reference's plugin API."""
from .attention import ATTENTION_REGISTRY, BaseAttention, FAVORPlusAttention, ReLUAttention, SoftmaxAttention
from .rpe import RPE_REGISTRY, BaseRPE, CirculantStringRPE, KERPLEPositionalEncoding, RoPE
from .factory import MODEL_VARIANTS, create_model, get_model_info, list_available_models
from .vit import BaseViT, UnifiedTransformerBlock
from .configs import CIFAR10_CONFIG, MNIST_CONFIG

__all__ = [
    "ATTENTION_REGISTRY", "RPE_REGISTRY", "BaseAttention", "SoftmaxAttention", "FAVORPlusAttention", "ReLUAttention",
    "BaseRPE", "KERPLEPositionalEncoding", "CirculantStringRPE", "RoPE", "MODEL_VARIANTS", "create_model",
    "get_model_info", "list_available_models", "BaseViT", "UnifiedTransformerBlock", "MNIST_CONFIG", "CIFAR10_CONFIG",
]
