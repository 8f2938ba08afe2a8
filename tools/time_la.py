import sys, torch
sys.path.insert(0,'efficient-rpe-vit_b200')
from erv_b200 import ops
torch.manual_seed(0)
B,N,H,DH,M=1024,65,2,16,256
qkv=torch.randn(B,N,3*H*DH,device='cuda',requires_grad=True)
omega=torch.randn(H,DH,M,device='cuda')
g=torch.randn(B,N,H*DH,device='cuda')
for save in (True, False):
    ops.SAVE_KV_STATE = save
    for _ in range(3):
        out=ops.linear_attention(qkv,omega,H,ops.FEAT_FAVOR); out.backward(g); qkv.grad=None
    torch.cuda.synchronize()
    e=[torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    for _ in range(10): out=ops.linear_attention(qkv,omega,H,ops.FEAT_FAVOR)
    e[1].record()
    for _ in range(10): out.backward(g, retain_graph=True); qkv.grad=None
    e[2].record()
    with torch.no_grad():
        for _ in range(10): o2=ops.linear_attention(qkv,omega,H,ops.FEAT_FAVOR)
    e[3].record(); torch.cuda.synchronize()
    print("save_state=%s fwd ms %.4f  bwd ms %.4f  fwd(no_grad) ms %.4f"%(save, e[0].elapsed_time(e[1])/10, e[1].elapsed_time(e[2])/10, e[2].elapsed_time(e[3])/10))
