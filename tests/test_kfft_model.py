"""CPU model of the KERPLE FFT route's algorithm (efficient-rpe-vit_b200/csrc/erv_kerple_fft.cu), in numpy:
  * the self-sorting 16 x 16 x 32 decomposition of the 8192-point DFT, phase by phase with the device kernel's index maps
    (thread t holds x[t + 512 r]; the result lands as X[t + 512 q]),
  * the inverse as conj(DFT(conj(.))) / L and two real columns per complex transform,
  * the CLS split: the patches' Toeplitz product by circular convolution of length 2 (N - 1) plus the rank-1 CLS row / column,
against numpy's FFT and the dense Toeplitz-masked attention the reference's route is algebraically equal to
(kerple.py:252-344, fft_utils.py:142-170, favor_plus.py:221-260).  No GPU needed."""
import numpy as np

L, NT = 8192, 512


def fft8192_model(x):
    """x[r, t] = element t + 512 r (r < 16, t < 512) -> X[q, t] = X[t + 512 q], exactly as fft8192() in the kernel."""
    t = np.arange(NT)
    # phase A: 16-point DFT over r, twiddle w_L^(t k1)
    y = np.fft.fft(x, axis=0) * np.exp(-2j * np.pi * np.outer(np.arange(16), t) / L)          # y[k1, t]
    # exchange 1: sm[k1 * 512 + t]; thread (k1 = warp, m2 = lane) reads sm[warp * 512 + 32 m1 + lane]
    sm = y.reshape(-1)
    warp, lane = t >> 5, t & 31
    xb = np.stack([sm[warp * 512 + 32 * m1 + lane] for m1 in range(16)])                      # xb[m1, thread]
    # phase B: 16-point DFT over m1, twiddle w_512^(lane j1)
    z = np.fft.fft(xb, axis=0) * np.exp(-2j * np.pi * np.outer(np.arange(16), lane) / 512)    # z[j1, thread]
    # exchange 2: sm[(j1 * 16 + warp) * 33 + lane]
    sm2 = np.zeros(256 * 33, dtype=complex)
    for j1 in range(16):
        sm2[(j1 * 16 + warp) * 33 + lane] = z[j1]
    # phase C: thread (k1 = t & 15, j1 = (t >> 4) & 15, half = t >> 8): 32-point DFT, outputs j2 = 2 q + half
    row = (((t >> 4) & 15) * 16 + (t & 15)) * 33
    a = np.stack([sm2[row + u] for u in range(16)])
    b = np.stack([sm2[row + 16 + u] for u in range(16)])
    odd = (t >= 256)[None, :]
    w32 = np.exp(-2j * np.pi * np.arange(16) / 32)[:, None]
    xc = np.where(odd, (a - b) * w32, a + b)
    return np.fft.fft(xc, axis=0)                                                              # X[q, t]


def test_three_phase_transform_is_the_dft_in_place():
    rng = np.random.default_rng(0)
    seq = rng.standard_normal(L) + 1j * rng.standard_normal(L)
    got = fft8192_model(seq.reshape(16, NT))          # element t + 512 r at [r, t]
    want = np.fft.fft(seq).reshape(16, NT)            # X[t + 512 q] at [q, t]
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()


def test_two_real_columns_per_transform_and_conjugate_inverse():
    rng = np.random.default_rng(1)
    n = 300                                            # patches; circular length L >= 2 n - 1
    c = np.exp(0.3 * rng.standard_normal(2 * n - 1))   # c[delta + n - 1], delta = j - i
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    g = np.zeros(L)
    g[:n] = c[n - 1 - np.arange(n)]                    # g[s] = c[-s]
    g[L - np.arange(1, n)] = c[n - 1 + np.arange(1, n)]  # g[L - s] = c[s]
    G = np.fft.fft(g) / L
    x = np.zeros(L, dtype=complex)
    x[:n] = a + 1j * b
    X = fft8192_model(x.reshape(16, NT)).reshape(-1)
    y = np.conj(fft8192_model(np.conj(X * G).reshape(16, NT)).reshape(-1))
    C = np.array([[c[j - i + n - 1] for j in range(n)] for i in range(n)])
    assert np.abs(y.real[:n] - C @ a).max() < 1e-9 * np.abs(C @ a).max()
    assert np.abs(y.imag[:n] - C @ b).max() < 1e-9 * np.abs(C @ b).max()


def test_cls_split_equals_dense_masked_attention():
    rng = np.random.default_rng(2)
    N, M, D = 41, 6, 4                                  # tokens (CLS + 40 patches), features, head_dim
    pq, pk = np.abs(rng.standard_normal((N, M))), np.abs(rng.standard_normal((N, M)))
    v = rng.standard_normal((N, D))
    c = np.exp(0.3 * rng.standard_normal(2 * N - 1))    # c[delta + N - 1]
    A = (pq @ pk.T) * np.array([[c[j - i + N - 1] for j in range(N)] for i in range(N)])
    u = np.concatenate([v, np.ones((N, 1))], axis=1)
    want = A @ u                                         # [num | den]
    # patches by circular convolution (any length >= 2 (N - 1) - 1), one (feature, column) pair at a time
    n, Lc = N - 1, 128
    g = np.zeros(Lc)
    g[:n] = c[N - 1 - np.arange(n)]
    g[Lc - np.arange(1, n)] = c[N - 1 + np.arange(1, n)]
    G = np.fft.fft(g)
    got = np.zeros((N, D + 1))
    for m in range(M):
        for d in range(D + 1):
            x = np.zeros(Lc)
            x[:n] = pk[1:, m] * u[1:, d]
            got[1:, d] += pq[1:, m] * np.fft.ifft(G * np.fft.fft(x)).real[:n]
    # CLS column (kfft_finalize_kernel) and CLS row (kfft_cls_kernel)
    s_col = (pq[1:] @ pk[0]) * c[N - 1 - np.arange(1, N)]          # c[0 - i]
    got[1:] += s_col[:, None] * u[0][None, :]
    s_row = (pk @ pq[0]) * c[N - 1 + np.arange(N)]                 # c[j - 0]
    got[0] = s_row @ u
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()
