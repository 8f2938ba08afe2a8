# per-launch GPU time of the short-sequence linear-attention kernels at BASELINE config 2: 10 C-ABI calls captured in a CUDA
# graph (no Python / launch overhead between kernels), replayed 5 times, CUDA events around each replay: median and best
import sys, statistics, torch
sys.path.insert(0,'efficient-rpe-vit_b200')
from erv_b200 import ops, _capi as C
lib=C.load()
torch.manual_seed(0)
B,N,H,DH,M=1024,65,2,16,256
if len(sys.argv) > 1: M = int(sys.argv[1])
kind=ops.FEAT_FAVOR
dt=torch.float32 if len(sys.argv) < 3 else torch.bfloat16
qkv=torch.randn(B,N,3*H*DH,device='cuda',dtype=dt)
omega=torch.randn(H,DH,M,device='cuda')
g=torch.randn(B,N,H*DH,device='cuda',dtype=dt)
out=torch.empty(B,N,H*DH,device='cuda',dtype=dt)
dqkv=torch.empty_like(qkv)
nb=lib.erv_linear_attention_workspace(B,N,H,DH,M,0,1)
ws=C.workspace(nb, qkv.device)
nstate=lib.erv_linear_attention_state_floats(B,N,H,DH,M)
state=torch.empty(max(nstate,1),device='cuda')
REP=10
def fwd(st):
    C.check(lib.erv_linear_attention_fwd(C.ptr(qkv),C.ptr(out),C.ptr(omega),B,N,H,DH,M,kind,0,None,None,C.dtype_code(qkv),C.ptr(st) if st is not None else None,C.ptr(ws),nb,C.stream()),"f")
def bwd(st):
    C.check(lib.erv_linear_attention_bwd(C.ptr(qkv),C.ptr(out),C.ptr(g),C.ptr(dqkv),C.ptr(omega),B,N,H,DH,M,kind,0,None,None,None,C.dtype_code(qkv),C.ptr(st) if st is not None else None,C.ptr(ws),nb,C.stream()),"b")
def graph_time(fn):
    s=torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr=torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(REP): fn()
    ts=[]
    for _ in range(6):
        torch.cuda.synchronize()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b)/REP)
    return statistics.median(ts[1:]), min(ts[1:])
st = state if nstate else None
print("M=%d %s  fwd+state ms %.4f (min %.4f)  fwd ms %.4f (min %.4f)  bwd(state) ms %.4f (min %.4f)  bwd(recompute) ms %.4f (min %.4f)" % (
    M, dt, *graph_time(lambda: fwd(st)), *graph_time(lambda: fwd(None)), *graph_time(lambda: bwd(st)), *graph_time(lambda: bwd(None))))
