"""Toeplitz products (reference: models/rpe/fft_utils.py).  `fft_toeplitz_matmul` keeps the reference's
name and shape rules; on the GPU it is a direct tiled product (no FFT, no 2n-1 padding)."""
import torch

from .. import ops


def fft_toeplitz_matmul(c: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """y = T x, T[i,j] = c[j-i+n-1].  c: [2n-1] with x [n] / [n,d] / [B,n,d], or c [B,H,2n-1] with
    x [B,H,n,d] (fft_utils.py:17-84)."""
    if c.dim() == 1:
        if x.dim() == 3:
            _check_n(c.shape[0], x.shape[1])
            return ops.toeplitz_matmul(c.unsqueeze(0), x)
        if x.dim() == 2:
            _check_n(c.shape[0], x.shape[0])
            return ops.toeplitz_matmul(c.unsqueeze(0), x.unsqueeze(0))[0]
        if x.dim() == 1:
            _check_n(c.shape[0], x.shape[0])
            return ops.toeplitz_matmul(c.unsqueeze(0), x.view(1, -1, 1))[0, :, 0]
        raise ValueError(f"x must have 1 or 2 dimensions. Got shape={x.shape}")
    if c.dim() == 3:
        if x.dim() != 4:
            raise ValueError(f"When c has 3 dims, x must have 4 dims. Got x.shape={x.shape}")
        b, h, n, d = x.shape
        assert c.shape[0] == b and c.shape[1] == h, "Batch and head dimensions must match"
        _check_n(c.shape[2], n)
        return ops.toeplitz_matmul(c.reshape(b * h, -1), x.reshape(b * h, n, d)).reshape(b, h, n, d)
    raise ValueError(f"c must have 1 or 3 dimensions. Got shape={c.shape}")


def _check_n(n_coeffs: int, rows: int):
    n = (n_coeffs + 1) // 2
    assert rows == n, f"Matrix height {rows} doesn't match expected {n} from coefficients {n_coeffs}"


def create_toeplitz_matrix(c: torch.Tensor, n: int) -> torch.Tensor:
    """Explicit T[i,j] = c[j-i+n-1] (fft_utils.py:261-292); for tests."""
    i = torch.arange(n, device=c.device)
    return c[(i[None, :] - i[:, None]) + (n - 1)]


def naive_toeplitz_matmul(c: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    return create_toeplitz_matrix(c, x.shape[0]) @ x
