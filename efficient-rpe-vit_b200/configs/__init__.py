"""Drop-in alias of the reference's `configs` package (configs/__init__.py:7-18) and its sub-modules (`configs.base`,
`configs.datasets.mnist`, `configs.datasets.cifar10`, `configs.mnist_config`, `configs.cifar10_config`)."""
import sys
import types

from erv_b200.configs import (CIFAR10_CONFIG, MNIST_CONFIG, BaseConfig, CIFAR10Config, MNISTConfig, get_attention_config,
                              get_rpe_config)

_mnist = {"MNIST_CONFIG": MNIST_CONFIG, "MNISTConfig": MNISTConfig}
_cifar = {"CIFAR10_CONFIG": CIFAR10_CONFIG, "CIFAR10Config": CIFAR10Config}
for _name, _attrs in {
    "base": {"BaseConfig": BaseConfig, "get_attention_config": get_attention_config, "get_rpe_config": get_rpe_config},
    "mnist_config": _mnist, "cifar10_config": _cifar, "datasets": {**_mnist, **_cifar},
    "datasets.mnist": {**_mnist, "BaseConfig": BaseConfig}, "datasets.cifar10": {**_cifar, "BaseConfig": BaseConfig},
}.items():
    _m = types.ModuleType(f"{__name__}.{_name}")
    _m.__dict__.update(_attrs)
    sys.modules[f"{__name__}.{_name}"] = _m

__all__ = ["BaseConfig", "MNISTConfig", "CIFAR10Config", "MNIST_CONFIG", "CIFAR10_CONFIG", "get_attention_config",
           "get_rpe_config"]
