"""ctypes binding of liberv_b200.so (the C ABI declared in include/erv_b200.h).

There is no fallback: if the library is missing or a tensor is not on a CUDA device the call
raises.  The library is built in-tree by ``efficient-rpe-vit_b200/csrc/build.py``.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p  # noqa: F401

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "liberv_b200.so")

OK, E_INVALID, E_UNSUPPORTED, E_CUDA, E_WORKSPACE = 0, 1, 2, 3, 4
F32, BF16 = 0, 1
FEAT_FAVOR, FEAT_RELU = 0, 1
ROT_NONE, ROT_ROPE, ROT_CIRCULANT = 0, 1, 2

# name -> (restype, argtypes); must list every symbol of include/erv_b200.h (tests/test_capi.py checks)
_P, _I, _F, _Z = c_void_p, c_int, c_float, c_size_t
SIGNATURES = {
    "erv_abi_version": (c_int, []),
    "erv_last_error": (c_char_p, []),
    "erv_launch_count": (c_uint64, []),
    "erv_reset_launch_count": (None, []),
    "erv_rope_table": (c_int, [_F, _I, _I, _P, _P, _P]),
    "erv_circulant_table_fwd": (c_int, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "erv_circulant_table_bwd_scratch": (c_size_t, [_I, _I, _I, _I]),
    "erv_circulant_table_bwd": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "erv_rotate": (c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "erv_rotate_table_grad": (c_int, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "erv_feature_map_workspace": (c_size_t, [_I, _I, _I]),
    "erv_feature_map_fwd": (c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "erv_feature_map_bwd": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "erv_linear_attention_workspace": (c_size_t, [_I, _I, _I, _I, _I, _I, _I]),
    "erv_kerple_attention_workspace": (c_size_t, [_I, _I, _I, _I, _I, _I]),
    "erv_softmax_attention_workspace": (c_size_t, [_I, _I, _I, _I, _I, _I]),
    "erv_circulant_slots": (c_int, [_I, _I]),
    "erv_linear_attention_fwd": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _Z, _P]),
    "erv_linear_attention_bwd": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _Z, _P]),
    "erv_linear_attention_state_floats": (c_size_t, [_I, _I, _I, _I, _I]),
    "erv_embed_supported": (c_int, [_I, _I]),
    "erv_embed_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "erv_embed_bwd_workspace": (c_size_t, [_I, _I, _I, _I]),
    "erv_embed_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "erv_head_loss_workspace": (c_size_t, [_I, _I]),
    "erv_head_loss_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _Z, _P]),
    "erv_head_loss_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _Z, _P]),
    "erv_block_supported": (c_int, [_I, _I]),
    "erv_block_set_tensor_core": (None, [_I]),
    "erv_block_ln_qkv_params": (c_int, []),
    "erv_block_mlp_params": (c_int, []),
    "erv_block_ln_qkv_bwd_workspace": (c_size_t, [_I]),
    "erv_block_mlp_bwd_workspace": (c_size_t, [_I]),
    "erv_block_act_bf16_supported": (c_int, []),
    "erv_block_ln_qkv_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "erv_block_ln_qkv_bwd": (c_int, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P, _Z, _P]),
    "erv_block_mlp_fwd": (c_int, [_P, _I, _P, _P, _P, _I, _I, _I, _F, _F, _P, _I, _P]),
    "erv_block_mlp_bwd": (c_int, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _P, _I, _P, _Z, _P]),
    "erv_kerple_set_fft": (None, [_I]),
    "erv_kerple_attention_fwd": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "erv_kerple_attention_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "erv_softmax_attention_fwd": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _F, c_uint64, _P, _I, _P, _Z, _P]),
    "erv_softmax_attention_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _F, c_uint64, _P, _I, _P, _Z, _P]),
    "erv_toeplitz_matmul_fwd": (c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "erv_toeplitz_matmul_bwd": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "erv_adam_step": (c_int, [_P, _P, _P, _P, _Z, _F, _F, _F, _F, _F, _I, _F, c_int64, _P, _P]),
    "erv_adam_step_dev": (c_int, [_P, _P, _P, _P, _Z, _P, _I, _F, _P, _P]),
    "erv_allreduce_flag_floats": (c_int, []),
    "erv_allreduce_oneshot": (c_int, [_P, _Z, _Z, _P, _I, _I, _P, _P]),
    "erv_linear_wgrad_supported": (c_int, [_I, _I, _I]),
    "erv_linear_wgrad_workspace": (c_size_t, [_I, _I, _I]),
    "erv_linear_wgrad": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    "erv_layernorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "erv_layernorm_bwd_workspace": (c_size_t, [_I, _I]),
    "erv_layernorm_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _Z, _P]),
    "erv_debug_umma_gemm": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "erv_debug_set_trace": (None, [_P]),
    "erv_debug_umma_timing": (c_int, [_I, _I, _I, _I, _I, _P, _P]),
    "erv_debug_umma_gemm_ts": (c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "erv_debug_umma_timing2": (c_int, [_I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "erv_debug_unit_probe": (c_int, [_I, _I, _I, _P, _P, _P]),
}

_lib = None


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"erv_b200: {_LIB_PATH} not found. Build it with "
            "`python efficient-rpe-vit_b200/csrc/build.py` (needs nvcc; sm_100a only). There is no CPU fallback.")
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class ErvError(RuntimeError):
    pass


def check(status: int, what: str = ""):
    if status == OK:
        return
    msg = load().erv_last_error().decode("utf-8", "replace")
    text = f"{what}: {msg}" if what else msg
    if status == E_INVALID:
        raise ValueError(text)
    if status == E_UNSUPPORTED:
        raise NotImplementedError(text)
    raise ErvError(f"{text} (status {status})")


def require_cuda(*tensors):
    """Every tensor must live on the CURRENT CUDA device: the C ABI launches on torch's current stream of the current
    device (stream()), and the per-function shared-memory attributes are cached per device."""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("erv_b200 runs on CUDA tensors only (sm_100a); got a tensor on "
                               f"'{t.device}'. There is no CPU fallback.")
        if t.device.index != torch.cuda.current_device():
            raise RuntimeError(f"erv_b200: tensor on {t.device} but the current device is cuda:{torch.cuda.current_device()}; "
                               "wrap the call in `with torch.cuda.device(tensor.device):`")


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"erv_b200 supports float32 and bfloat16 activations, got {t.dtype}")


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def launch_count() -> int:
    return int(load().erv_launch_count())


def reset_launch_count():
    load().erv_reset_launch_count()
