# needs a library built with the trace points: ERV_TRACE=1 python -c "import sys; sys.path.insert(0, 'efficient-rpe-vit_b200/csrc'); import build; build.build(force=True)"
import sys, torch, collections
sys.path.insert(0,'efficient-rpe-vit_b200')
from erv_b200 import ops, _capi as C
lib=C.load()
torch.manual_seed(0)
B,N,H,DH,M=1024,65,2,16,256
qkv=torch.randn(B,N,3*H*DH,device='cuda',requires_grad=True)
omega=torch.randn(H,DH,M,device='cuda')
out=ops.linear_attention(qkv,omega,H,ops.FEAT_FAVOR)
g=torch.randn_like(out)
out.backward(g, retain_graph=True)   # warm
torch.cuda.synchronize()
buf=torch.zeros(2000,dtype=torch.int64,device='cuda')
lib.erv_debug_set_trace(C.ptr(buf))
qkv.grad=None
out.backward(g)
torch.cuda.synchronize()
lib.erv_debug_set_trace(None)
t=buf.cpu().tolist()
ev=[(t[2*i],t[2*i+1]) for i in range(1000) if t[2*i+1]]
print(len(ev),'events; total cycles', ev[-1][1]-ev[0][1], 'tiles', sum(1 for e in ev if e[0]%100==0))
agg=collections.defaultdict(list)
for (a,ta),(b,tb) in zip(ev,ev[1:]):
    agg[(a,b)].append(tb-ta)
for k,v in sorted(agg.items()):
    print(k, 'n=%d avg=%.0f min=%d max=%d'%(len(v),sum(v)/len(v),min(v),max(v)))
