"""Drop-in alias: exposes erv_b200 under the reference's module paths (`models`, `models.attention`,
`models.rpe`, `models.rpe.fft_utils`, `models.factory`, `models.core.base_vit`,
`models.components.unified_transformer`), so code written against the reference imports unchanged when
this directory precedes the reference on sys.path."""
import sys
import types

import erv_b200
from erv_b200 import attention, factory, rpe, vit
from erv_b200.attention import base as _abase, favor_plus as _afavor, relu as _arelu, softmax as _asoftmax
from erv_b200.rpe import base as _rbase, circulant_string as _rcirc, fft_utils as _rfft, kerple as _rkerple, rope as _rrope
from erv_b200 import *  # noqa: F401,F403

_core = types.ModuleType(__name__ + ".core")
_components = types.ModuleType(__name__ + ".components")
_core.base_vit = vit
_core.BaseViT = vit.BaseViT
_components.unified_transformer = vit
_components.UnifiedTransformerBlock = vit.UnifiedTransformerBlock

for _name, _mod in {
    "attention": attention, "attention.base": _abase, "attention.softmax": _asoftmax,
    "attention.favor_plus": _afavor, "attention.relu": _arelu,
    "rpe": rpe, "rpe.base": _rbase, "rpe.rope": _rrope, "rpe.circulant_string": _rcirc, "rpe.kerple": _rkerple,
    "rpe.fft_utils": _rfft, "factory": factory, "core": _core, "core.base_vit": vit,
    "components": _components, "components.unified_transformer": vit,
}.items():
    sys.modules[f"{__name__}.{_name}"] = _mod

__all__ = erv_b200.__all__
