import time, torch, sys
sys.path.insert(0,'efficient-rpe-vit_b200')
from erv_b200 import ops
n = 1024*2*17*256
torch.cuda.synchronize()
keep = None
t0=time.perf_counter()
for i in range(20):
    new = torch.empty(n, device='cuda'); keep = new
torch.cuda.synchronize(); print('alloc loop ms/iter', (time.perf_counter()-t0)/20*1e3)
qkv=torch.randn(1024,65,96,device='cuda',requires_grad=True); omega=torch.randn(2,16,256,device='cuda')
for _ in range(3): out=ops.linear_attention(qkv,omega,2,ops.FEAT_FAVOR)
torch.cuda.synchronize(); t0=time.perf_counter()
for i in range(20): out=ops.linear_attention(qkv,omega,2,ops.FEAT_FAVOR)
torch.cuda.synchronize(); print('fwd loop (grad) ms/iter', (time.perf_counter()-t0)/20*1e3)
print(torch.cuda.memory_stats()['num_alloc_retries'], torch.cuda.memory_stats()['segment.all.allocated'])
import torch.profiler as P
with P.profile(activities=[P.ProfilerActivity.CPU, P.ProfilerActivity.CUDA]) as prof:
    for i in range(5): out=ops.linear_attention(qkv,omega,2,ops.FEAT_FAVOR)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8))
