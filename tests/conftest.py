import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-rpe-vit_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def golden_files(prefix):
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a - b| / max |b|: the worst element against the reference's scale (SURVEY.md 7 step 0 asks for rel-L2 AND a
    max-type check; a plain element-wise relative error is meaningless at the zero crossings of an attention output)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# max-type tolerance = MAX_FACTOR x the rel-L2 tolerance: a handful of ill-conditioned elements (normalisers close to the
# +1e-6 of favor_plus.py:260) carry several times the rms error while still being far inside fp32 round-off of the reference
MAX_FACTOR = 4.0


def assert_close(a, b, tol, what=""):
    e2, em = rel_l2(a, b), max_rel(a, b)
    assert e2 < tol, f"{what}: rel-L2 {e2:.3e} >= {tol:.1e}"
    assert em < MAX_FACTOR * tol, f"{what}: max-rel {em:.3e} >= {MAX_FACTOR * tol:.1e} (rel-L2 {e2:.3e})"


ATTN_KIND = {"softmax": "softmax", "favor_plus": "favor", "relu": "relu"}
RPE_KIND = {"none": None, "rope": "rope", "circulant_string": "circulant", "most_general": "kerple"}


def parse_attn_case(fname):
    """attn_<shape>_<attention>_<rpe>.npz -> (shape, attention registry name, rpe registry name or None)."""
    stem = fname[:-4].split("_", 2)
    shape, rest = stem[1], stem[2]
    for a in ("favor_plus", "softmax", "relu"):
        if rest.startswith(a + "_"):
            r = rest[len(a) + 1:]
            return shape, a, (None if r == "none" else r)
    raise ValueError(fname)
