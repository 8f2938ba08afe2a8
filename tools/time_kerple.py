# GPU time of the KERPLE (Toeplitz-masked) tile route, forward and backward, at the config-5 sequence length for M = 44 / 256
# and B = 2 / 8: CUDA events around 10 back-to-back calls after 3 warm-up calls (inputs ~8 MB: L2-resident, as in the step)
import sys, torch
sys.path.insert(0, 'efficient-rpe-vit_b200')
from erv_b200 import ops, _capi
lib = _capi.load()
H, DH = 2, 16
def run(B, N, M, kind):
    torch.manual_seed(0)
    qkv = torch.randn(B, N, 3 * H * DH, device='cuda', requires_grad=True)
    omega = torch.randn(H, DH, M, device='cuda')
    bias = (0.02 * torch.randn(H, 2 * N - 1, device='cuda')).requires_grad_()
    g = torch.randn(B, N, H * DH, device='cuda')
    def ev(): return torch.cuda.Event(enable_timing=True)
    tf, tb = [], []
    for i in range(13):
        a, b, c = ev(), ev(), ev()
        a.record(); o = ops.kerple_attention(qkv, omega, bias, H, kind); b.record(); o.backward(g); c.record()
        torch.cuda.synchronize()
        if i >= 3: tf.append(a.elapsed_time(b)); tb.append(b.elapsed_time(c))
        qkv.grad = None; bias.grad = None
    pairs = B * H
    print("B=%d N=%d M=%d kind=%d  fwd %.3f ms (%.1f us/pair)  bwd %.3f ms (%.1f us/pair)" % (
        B, N, M, kind, min(tf), 1e3 * min(tf) / pairs, min(tb), 1e3 * min(tb) / pairs), flush=True)
for mode in (0, 1):
    lib.erv_kerple_set_fft(mode)
    print("forward route:", "FFT (erv_kerple_fft.cu)" if mode else "Toeplitz-masked tiles (erv_ktile_tc.cu / erv_tileattn.cu)", flush=True)
    for (B, N, M) in [(2, 4097, 44), (8, 4097, 44), (32, 4097, 44), (2, 4097, 256), (8, 4097, 256), (8, 2049, 44), (64, 1025, 44), (128, 513, 44)]:
        run(B, N, M, ops.FEAT_FAVOR)
