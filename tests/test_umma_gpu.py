"""Pins the tcgen05 descriptor encodings (erv_umma.cuh) with a plain GEMM probe against torch."""
import itertools

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bf16,a_mn,b_mn", list(itertools.product([0, 1], [0, 1], [0, 1])))
@pytest.mark.parametrize("n,k", [(32, 16), (256, 16), (48, 64), (128, 128)])
def test_umma_probe_gemm(bf16, a_mn, b_mn, n, k):
    from erv_b200 import _capi as C
    if not bf16 and (a_mn or b_mn):
        pytest.skip("TF32 operands are only used K-major (MN-major TF32 needs the 128B_BASE32B swizzle; measured: "
                    "no-swizzle MN-major TF32 returns zeros on B200)")
    torch.manual_seed(n * 1000 + k + bf16)
    a = torch.randn(128, k, device="cuda")
    b = torch.randn(n, k, device="cuda")
    d = torch.full((128, n), float("nan"), device="cuda")
    C.check(C.load().erv_debug_umma_gemm(C.ptr(a), C.ptr(b), C.ptr(d), n, k, a_mn, b_mn, bf16, C.stream()), "probe")
    torch.cuda.synchronize()
    if bf16:
        ref = a.bfloat16().float() @ b.bfloat16().float().T
        tol = 1e-5
    else:
        ref = (a.double() @ b.double().T).float()
        tol = 2e-3  # tf32 operands (10-bit mantissa)
    err = float((d - ref).norm() / ref.norm())
    assert err < tol, err


@pytest.mark.parametrize("bf16,b_mn", [(1, 0), (1, 1), (0, 0)])
@pytest.mark.parametrize("n,k", [(32, 16), (64, 64), (256, 128), (48, 256)])
def test_umma_probe_gemm_a_in_tmem(bf16, b_mn, n, k):
    """A operand written to tensor memory with tcgen05.st (lane = row, K packed along columns)."""
    from erv_b200 import _capi as C
    if not bf16 and k > 128:
        pytest.skip("tf32 probe holds K <= 128 columns")
    torch.manual_seed(n * 1000 + k + bf16)
    a = torch.randn(128, k, device="cuda")
    b = torch.randn(n, k, device="cuda")
    d = torch.full((128, n), float("nan"), device="cuda")
    C.check(C.load().erv_debug_umma_gemm_ts(C.ptr(a), C.ptr(b), C.ptr(d), n, k, b_mn, bf16, C.stream()), "probe")
    torch.cuda.synchronize()
    if bf16:
        ref = a.bfloat16().float() @ b.bfloat16().float().T
        tol = 1e-5
    else:
        ref = (a.double() @ b.double().T).float()
        tol = 2e-3
    err = float((d - ref).norm() / ref.norm())
    assert err < tol, err
