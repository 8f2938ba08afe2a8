"""CPU checks of the host-side mirror of the reference's plugin API: registries, factory names, constructor
arguments, state_dict keys, error conventions.  Nothing here launches a kernel."""
import warnings

import pytest
import torch

import erv_b200
from erv_b200 import (ATTENTION_REGISTRY, CIFAR10_CONFIG, MNIST_CONFIG, MODEL_VARIANTS, RPE_REGISTRY,
                      CirculantStringRPE, FAVORPlusAttention, KERPLEPositionalEncoding, ReLUAttention, RoPE,
                      SoftmaxAttention, create_model, get_model_info, list_available_models)
from conftest import golden_files, load_golden


def test_registries_and_aliases():
    assert set(ATTENTION_REGISTRY) == {"softmax", "baseline", "favor_plus", "favor+", "performer", "relu"}
    assert set(RPE_REGISTRY) == {"most_general", "kerple", "circulant_string", "circulant", "rope", "rotary"}
    assert ATTENTION_REGISTRY["performer"] is FAVORPlusAttention and RPE_REGISTRY["kerple"] is KERPLEPositionalEncoding
    assert len(list_available_models()) == len(MODEL_VARIANTS) == 14
    info = get_model_info("performer_relu_most_general")
    assert info["attention_type"] == "relu" and info["rpe_type"] == "most_general" and info["has_rpe"]
    with pytest.raises(ValueError):
        get_model_info("nope")
    with pytest.raises(ValueError):
        create_model("nope_x", MNIST_CONFIG)


def test_reference_import_paths():
    from models.rpe.fft_utils import fft_toeplitz_matmul, create_toeplitz_matrix  # noqa: F401
    from models.attention import FAVORPlusAttention as F2
    from models.rpe import KERPLEPositionalEncoding as K2
    from models.core.base_vit import BaseViT  # noqa: F401
    from configs.mnist_config import MNIST_CONFIG as M2
    assert F2 is FAVORPlusAttention and K2 is KERPLEPositionalEncoding and M2 is MNIST_CONFIG


@pytest.mark.parametrize("fname", golden_files("model_"))
def test_state_dict_interchange_with_reference(fname):
    """Reference state_dicts load with strict=True: same keys, shapes, dtypes."""
    g = load_golden(fname)
    name = fname[len("model_"):-4]
    if name.startswith("cifar_"):
        model = create_model("performer_favor", CIFAR10_CONFIG, attention_config={"num_features": 256}, patch_size=4)
    else:
        model = create_model(name, MNIST_CONFIG)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(sd, strict=True)
    assert model.model_name and model.attention_type in ("softmax", "favor_plus", "relu")


def test_factory_does_not_mutate_config_and_accepts_overrides():
    before = repr(MNIST_CONFIG)
    m = create_model("performer_favor", MNIST_CONFIG, attention_config={"num_features": 64}, patch_size=4, dropout=0.0)
    assert m.transformer_blocks[0].attention.num_features == 64 and m.num_patches == 49
    assert repr(MNIST_CONFIG) == before
    n = sum(p.numel() for p in create_model("baseline", MNIST_CONFIG).parameters())
    assert n == 27914
    assert sum(p.numel() for p in create_model("performer_relu_most_general", MNIST_CONFIG).parameters()) == 28112


def test_attention_constructor_contract():
    a = FAVORPlusAttention(dim=32, heads=2)
    assert a.num_features == 44 and a.omega.shape == (2, 16, 44) and a.favor_scale == 16 ** -0.25
    assert a.redraw_counter.dtype == torch.int64 and a.scale == 16 ** -0.5
    r = ReLUAttention(dim=64, heads=4, num_features=8, use_orthogonal=True)
    w = r.omega[0]
    assert torch.allclose(w.T @ w, 16 * torch.eye(8), atol=1e-4)  # M <= Dh: orthogonal columns of norm sqrt(Dh)
    assert r.relu_scale == a.favor_scale
    with pytest.raises(AssertionError):
        SoftmaxAttention(dim=30, heads=4)
    assert set(SoftmaxAttention(32, 2).state_dict()) == {"qkv.weight", "proj.weight", "proj.bias"}
    assert set(a.state_dict()) == {"omega", "redraw_counter", "qkv.weight", "proj.weight", "proj.bias"}


def test_rpe_constructor_contract():
    rope = RoPE(num_patches=17, dim=32, heads=2, theta=10000.0, junk=1)
    assert rope.cos_cached.shape == (17, 8) and not rope.state_dict()  # non-persistent buffers
    circ = CirculantStringRPE(num_patches=17, dim=32, heads=2, image_size=28, patch_size=7)
    assert circ.circulant_coeffs.shape == (2, 2, 16) and circ.patch_positions.shape == (16, 2)
    assert circ.circulant_coeffs.abs().max() < 0.1
    assert circ.patch_positions[5].tolist() == [1.0, 1.0] and circ.patch_positions[3].tolist() == [3.0, 0.0]
    ev = circ.get_eigenvalues()
    assert ev.real.abs().max() < 1e-6
    with pytest.raises(ValueError):
        CirculantStringRPE(num_patches=18, dim=32, heads=2)
    with pytest.raises(ValueError):
        CirculantStringRPE(num_patches=17, dim=32, heads=2, block_size=5)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        CirculantStringRPE(num_patches=17, dim=32, heads=2, block_size=8)
    assert len(w) == 1 and "block-circulant" in str(w[0].message)
    only_cls = CirculantStringRPE(num_patches=1, dim=32, heads=2)
    q = torch.randn(1, 2, 1, 16)
    assert only_cls.apply_circulant_string(q, q)[0] is q
    k = KERPLEPositionalEncoding(num_patches=17, dim=32, heads=2)
    assert k.rel_pos_bias.shape == (2, 33)
    with pytest.raises(NotImplementedError):
        k(torch.zeros(1))
    assert rope(q) is q and circ(q) is q


def test_error_conventions_without_gpu():
    x = torch.randn(2, 17, 32)
    sm = SoftmaxAttention(32, 2)
    with pytest.raises(NotImplementedError) as e:
        sm(x, rpe=KERPLEPositionalEncoding(17, 32, 2))
    assert "KERPLE" in str(e.value) and "kernelized" in str(e.value)
    with pytest.raises(AssertionError):  # RoPE: N <= num_patches (rope.py:93)
        FAVORPlusAttention(32, 2)(x, rpe=RoPE(num_patches=9, dim=32, heads=2))
    with pytest.raises(AssertionError):  # KERPLE: N == num_patches (fft_utils.py:138)
        ReLUAttention(32, 2)(x, rpe=KERPLEPositionalEncoding(9, 32, 2))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            sm(x)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            FAVORPlusAttention(32, 2)(x)


def test_product_never_imports_oracle():
    import os
    root = os.path.dirname(erv_b200.__file__)
    for dp, _, files in os.walk(os.path.dirname(root)):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dp, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "erv_oracle" not in text, f
