# phase trace of the pipelined forward / backward; needs the kernel compiled with -DERV_TRACE (tools/r2_trace_pipe.sh)
#   python tools/trace_pipe.py fwd|bwd
import sys, torch, collections
sys.path.insert(0,'efficient-rpe-vit_b200')
from erv_b200 import ops, _capi as C
lib=C.load()
which = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
torch.manual_seed(0)
B,N,H,DH,M=1024,65,2,16,256
qkv=torch.randn(B,N,3*H*DH,device='cuda',requires_grad=True)
omega=torch.randn(H,DH,M,device='cuda')
g=torch.randn(B,N,H*DH,device='cuda')
out=ops.linear_attention(qkv,omega,H,ops.FEAT_FAVOR)
out.backward(g, retain_graph=True)
torch.cuda.synchronize()
buf=torch.zeros(3000,dtype=torch.int64,device='cuda')
if which == 'fwd':
    lib.erv_debug_set_trace(C.ptr(buf))
    out2=ops.linear_attention(qkv,omega,H,ops.FEAT_FAVOR)
else:
    qkv.grad=None
    lib.erv_debug_set_trace(C.ptr(buf))
    out.backward(g)
torch.cuda.synchronize()
lib.erv_debug_set_trace(None)
t=buf.cpu().tolist()
for seg,name in enumerate(('compute warp 0','lone-token warp 17','issue warp')):
    s=t[seg*1000:(seg+1)*1000]
    ev=[(s[2*i],s[2*i+1]) for i in range(500) if s[2*i+1]]
    if not ev: continue
    print(name, len(ev),'events; total cycles', ev[-1][1]-ev[0][1])
    agg=collections.defaultdict(list)
    for (a,ta),(b,tb) in zip(ev,ev[1:]):
        agg[(a,b)].append(tb-ta)
    for k,v in sorted(agg.items()):
        print('  ',k, 'n=%d avg=%.0f min=%d max=%d'%(len(v),sum(v)/len(v),min(v),max(v)))
