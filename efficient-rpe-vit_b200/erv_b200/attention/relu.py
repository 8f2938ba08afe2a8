"""Performer-ReLU plugin (reference: models/attention/relu.py)."""
from .. import ops
from ._linear import RandomFeatureAttention


class ReLUAttention(RandomFeatureAttention):
    """phi(x) = relu(x W) / sqrt(M)."""
    _kind = ops.FEAT_RELU

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.relu_scale = self.head_dim ** -0.25

    def _compute_relu_features(self, x, omega):  # relu.py:116-138
        return self._features(x, omega)
