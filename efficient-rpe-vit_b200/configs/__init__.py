"""Drop-in alias of the reference's `configs` package (configs/__init__.py:7-18)."""
import sys
import types

from erv_b200.configs import CIFAR10_CONFIG, MNIST_CONFIG, get_attention_config, get_rpe_config

for _name, _attrs in {
    "mnist_config": {"MNIST_CONFIG": MNIST_CONFIG}, "cifar10_config": {"CIFAR10_CONFIG": CIFAR10_CONFIG},
    "datasets": {}, "datasets.mnist": {"MNIST_CONFIG": MNIST_CONFIG}, "datasets.cifar10": {"CIFAR10_CONFIG": CIFAR10_CONFIG},
}.items():
    _m = types.ModuleType(f"{__name__}.{_name}")
    _m.__dict__.update(_attrs)
    sys.modules[f"{__name__}.{_name}"] = _m

__all__ = ["MNIST_CONFIG", "CIFAR10_CONFIG", "get_attention_config", "get_rpe_config"]
