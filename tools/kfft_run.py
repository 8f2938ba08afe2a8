# one KERPLE forward on the FFT route at N = 4097 (driver for ncu: -k regex:kfft_fwd_kernel); usage: kfft_run.py [B] [M]
import sys, torch
sys.path.insert(0, 'efficient-rpe-vit_b200')
from erv_b200 import ops, _capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = int(sys.argv[2]) if len(sys.argv) > 2 else 256
H, DH, N = 2, 16, 4097
_capi.load().erv_kerple_set_fft(1)
torch.manual_seed(0)
qkv = torch.randn(B, N, 3 * H * DH, device='cuda')
omega = torch.randn(H, DH, M, device='cuda')
bias = 0.02 * torch.randn(H, 2 * N - 1, device='cuda')
for _ in range(3):
    o = ops.kerple_attention(qkv, omega, bias, H, ops.FEAT_FAVOR)
torch.cuda.synchronize()
print("ok", float(o.abs().mean()))
