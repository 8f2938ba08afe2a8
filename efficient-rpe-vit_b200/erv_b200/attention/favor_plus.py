"""FAVOR+ plugin (reference: models/attention/favor_plus.py)."""
from .. import ops
from ._linear import RandomFeatureAttention


class FAVORPlusAttention(RandomFeatureAttention):
    """phi(x) = exp(x W - max_f(x W) - |x|^2/2) / sqrt(M); the max is a constant for autograd."""
    _kind = ops.FEAT_FAVOR

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.favor_scale = self.head_dim ** -0.25

    def _compute_phi_positive(self, x, omega):  # favor_plus.py:112-140
        return self._features(x, omega)
