"""Dataset configs as plain dicts (reference: configs/base.py, configs/datasets/{mnist,cifar10}.py)."""
import copy

_ATTENTION_PARAMS = {
    "softmax": {},
    "favor_plus": {"num_features": None, "use_orthogonal": True, "feature_redraw_interval": None},
    "relu": {},
}
_RPE_PARAMS = {"most_general": {}, "circulant_string": {}, "rope": {"theta": 10000.0}}

_COMMON = dict(dim=32, depth=3, heads=2, mlp_dim=64, dropout=0.1, learning_rate=0.001, warmup_epochs=0,
               augmentation=False, num_workers=2, pin_memory=True, seed=42)

_MNIST = dict(_COMMON, image_size=28, in_channels=1, patch_size=7, num_classes=10, batch_size=32, weight_decay=0.0,
              epochs=10, mean=(0.1307,), std=(0.3081,))
_CIFAR10 = dict(_COMMON, image_size=32, in_channels=3, patch_size=8, num_classes=10, batch_size=64, weight_decay=0.01,
                epochs=20, warmup_epochs=2, mean=(0.4914, 0.4822, 0.4465), std=(0.2470, 0.2435, 0.2616))


def _with_params(cfg):
    out = dict(cfg)
    out["attention_params"] = copy.deepcopy(_ATTENTION_PARAMS)
    out["rpe_params"] = copy.deepcopy(_RPE_PARAMS)
    return out


MNIST_CONFIG = _with_params(_MNIST)
CIFAR10_CONFIG = _with_params(_CIFAR10)


def get_attention_config(attention_type: str, config=None):
    return dict((config or MNIST_CONFIG).get("attention_params", {}).get(attention_type, {}))


def get_rpe_config(rpe_type: str, config=None):
    return dict((config or MNIST_CONFIG).get("rpe_params", {}).get(rpe_type, {}))
