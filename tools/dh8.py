import sys, torch
sys.path.insert(0,'efficient-rpe-vit_b200'); sys.path.insert(0,'.')
from erv_b200 import ops, _capi
torch.manual_seed(0)
for (B,N,H,DH,M,rot) in [(3,50,8,8,24,0),(3,50,8,8,24,2),(2,65,2,8,24,0),(3,50,8,16,24,0),(1,50,1,8,16,0)]:
    qkv = torch.randn(B,N,3*H*DH, device='cuda')
    omega = torch.randn(H,DH,M, device='cuda')
    gt = None
    if rot==2:
        g = torch.zeros(H,N,DH, device='cuda'); g[:,:,0]=1; gt=g
    try:
        out = ops.linear_attention(qkv, omega, H, ops.FEAT_RELU, rot, gt)
        torch.cuda.synchronize()
        print((B,N,H,DH,M,rot), 'ok', float(out.abs().mean()))
    except Exception as e:
        print((B,N,H,DH,M,rot), 'FAIL', str(e)[:100]); break
