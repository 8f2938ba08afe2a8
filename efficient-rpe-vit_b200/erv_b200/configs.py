"""Dataset configurations (reference: configs/base.py:10-83, configs/datasets/mnist.py:10-52, configs/datasets/cifar10.py:10-51).

The reference exposes each configuration twice: as a class of UPPER_CASE attributes (`MNISTConfig`, `CIFAR10Config`, both
deriving from `BaseConfig` with `to_dict()` / `update(**kw)` class methods) and as the lower-case dict `to_dict()` returns
(`MNIST_CONFIG`, `CIFAR10_CONFIG`).  Both spellings are kept; here the values live in one table per dataset and the classes
are generated from it.
"""
import copy
from typing import Any, Dict

_ATTENTION_PARAMS = {
    "softmax": {},
    "favor_plus": {"num_features": None, "use_orthogonal": True, "feature_redraw_interval": None},
    "relu": {},
}
_RPE_PARAMS = {"most_general": {}, "circulant_string": {}, "rope": {"theta": 10000.0}}

# configs/base.py:13-42 (None = "must be set by the dataset config"; such entries are dropped by to_dict())
_BASE = dict(image_size=None, in_channels=None, patch_size=None, num_classes=None, dim=64, depth=3, heads=4, mlp_dim=256,
             dropout=0.1, batch_size=32, learning_rate=0.001, weight_decay=0.0, epochs=10, warmup_epochs=0, mean=None,
             std=None, augmentation=False, num_workers=2, pin_memory=True, seed=42)
_MODEL = dict(dim=32, depth=3, heads=2, mlp_dim=64, dropout=0.1)
_MNIST = dict(_BASE, **_MODEL, image_size=28, in_channels=1, patch_size=7, num_classes=10, batch_size=32,
              mean=(0.1307,), std=(0.3081,), num_workers=0)
_CIFAR10 = dict(_BASE, **_MODEL, image_size=32, in_channels=3, patch_size=8, num_classes=10, batch_size=64,
                weight_decay=0.01, epochs=20, warmup_epochs=2, mean=(0.4914, 0.4822, 0.4465),
                std=(0.2470, 0.2435, 0.2616))


class _ConfigMeta(type):
    """Builds the UPPER_CASE class attributes from a `_values` table."""

    def __new__(mcs, name, bases, ns):
        for k, v in ns.get("_values", {}).items():
            ns.setdefault(k.upper(), v)
        return super().__new__(mcs, name, bases, ns)


class BaseConfig(metaclass=_ConfigMeta):
    _values = dict(_BASE, attention_params=_ATTENTION_PARAMS, rpe_params=_RPE_PARAMS)

    @classmethod
    def to_dict(cls) -> Dict[str, Any]:
        """Lower-case dict of every public UPPER_CASE attribute that is not None (configs/base.py:64-73).  Nested parameter
        tables are deep-copied: the factory mutates them."""
        out = {}
        for key in dir(cls):
            if key.isupper() and not key.startswith("_"):
                value = getattr(cls, key)
                if value is not None:
                    out[key.lower()] = copy.deepcopy(value) if isinstance(value, dict) else value
        return out

    @classmethod
    def update(cls, **kwargs) -> Dict[str, Any]:
        out = cls.to_dict()
        out.update(kwargs)
        return out


class MNISTConfig(BaseConfig):
    _values = _MNIST


class CIFAR10Config(BaseConfig):
    _values = _CIFAR10


MNIST_CONFIG = MNISTConfig.to_dict()
CIFAR10_CONFIG = CIFAR10Config.to_dict()


def get_attention_config(attention_type: str, base_config=None) -> Dict[str, Any]:
    """configs/base.py:86-99; also accepts a config dict (or nothing: MNIST)."""
    return dict(_params(base_config, "attention_params").get(attention_type, {}))


def get_rpe_config(rpe_type: str, base_config=None) -> Dict[str, Any]:
    """configs/base.py:102-115."""
    return dict(_params(base_config, "rpe_params").get(rpe_type, {}))


def _params(base_config, key):
    if base_config is None:
        return MNIST_CONFIG.get(key, {})
    if isinstance(base_config, dict):
        return base_config.get(key, {})
    return getattr(base_config, key.upper(), {})
