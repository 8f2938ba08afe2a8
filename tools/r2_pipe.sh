#!/bin/bash
# pipelined short-sequence forward: parity against the previous kernel and the golden files, timing
tag=${1:-p}
timeout 120 python - > gpurun_out/r2_${tag}_cmp.txt 2>&1 <<'PY'
import os, sys, subprocess, torch
sys.path.insert(0, 'efficient-rpe-vit_b200')
from erv_b200 import ops
def run(B, N, M, kind, seed=0, dtype=torch.float32):
    torch.manual_seed(seed)
    H, DH = 2, 16
    qkv = torch.randn(B, N, 3 * H * DH, device='cuda', dtype=dtype)
    omega = torch.randn(H, DH, M, device='cuda')
    with torch.no_grad():
        return ops.linear_attention(qkv, omega, H, kind)
code = r'''
import sys, torch
sys.path.insert(0, 'efficient-rpe-vit_b200')
sys.path.insert(0, 'tools')
from erv_b200 import ops
H, DH = 2, 16
res = {}
for (B, N, M, kind) in [(4, 65, 256, ops.FEAT_FAVOR), (5, 65, 200, ops.FEAT_FAVOR), (3, 64, 256, ops.FEAT_RELU), (7, 50, 256, ops.FEAT_FAVOR), (1024, 65, 256, ops.FEAT_FAVOR), (33, 37, 129, ops.FEAT_RELU), (301, 65, 256, ops.FEAT_FAVOR)]:
    torch.manual_seed(1)
    qkv = torch.randn(B, N, 3 * H * DH, device='cuda')
    omega = torch.randn(H, DH, M, device='cuda')
    with torch.no_grad():
        o = ops.linear_attention(qkv, omega, H, kind)
    torch.cuda.synchronize()
    res[(B, N, M, kind)] = o.cpu()
torch.save(res, sys.argv[1])
'''
open('/tmp/cmp_run.py', 'w').write(code)
for name, env in (('new', {}), ('old', {'ERV_DISABLE_PIPE': '1'})):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, '/tmp/cmp_run.py', '/tmp/cmp_%s.pt' % name], env=e, capture_output=True, text=True, timeout=100)
    print(name, 'rc', r.returncode, r.stderr[-1500:])
a, b = torch.load('/tmp/cmp_new.pt'), torch.load('/tmp/cmp_old.pt')
for k in a:
    d = (a[k] - b[k]).norm() / b[k].norm()
    print(k, 'rel_l2 %.3e max_abs %.3e nan %d' % (d, (a[k] - b[k]).abs().max(), int(torch.isnan(a[k]).sum())))
PY
cat gpurun_out/r2_${tag}_cmp.txt
timeout 300 python -m pytest tests/test_parity_gpu.py -q -x -k "favor or relu or linear" 2>&1 | tail -8 > gpurun_out/r2_${tag}_tests.txt
cat gpurun_out/r2_${tag}_tests.txt
timeout 120 python tools/time_la.py 2>&1 | tee gpurun_out/r2_${tag}_time.txt
