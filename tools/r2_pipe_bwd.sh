#!/bin/bash
# pipelined short-sequence backward: parity against the previous kernel and the golden files, timing
tag=${1:-pb}
timeout 200 python - > gpurun_out/r2_${tag}_cmp.txt 2>&1 <<'PY'
import os, sys, subprocess, torch
code = r'''
import sys, torch
sys.path.insert(0, 'efficient-rpe-vit_b200')
from erv_b200 import ops
H, DH = 2, 16
res = {}
for (B, N, M, kind) in [(4, 65, 256, ops.FEAT_FAVOR), (5, 65, 200, ops.FEAT_FAVOR), (3, 64, 256, ops.FEAT_RELU), (7, 50, 256, ops.FEAT_FAVOR), (1024, 65, 256, ops.FEAT_FAVOR), (33, 37, 129, ops.FEAT_RELU), (301, 65, 256, ops.FEAT_FAVOR), (2, 65, 256, ops.FEAT_RELU)]:
    torch.manual_seed(1)
    qkv = torch.randn(B, N, 3 * H * DH, device='cuda', requires_grad=True)
    omega = torch.randn(H, DH, M, device='cuda')
    g = torch.randn(B, N, H * DH, device='cuda')
    o = ops.linear_attention(qkv, omega, H, kind)
    o.backward(g)
    torch.cuda.synchronize()
    res[(B, N, M, kind)] = (o.detach().cpu(), qkv.grad.cpu())
torch.save(res, sys.argv[1])
'''
open('/tmp/cmp_run.py', 'w').write(code)
for name, env in (('new', {}), ('old', {'ERV_DISABLE_PIPE_BWD': '1'})):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, '/tmp/cmp_run.py', '/tmp/cmp_%s.pt' % name], env=e, capture_output=True, text=True, timeout=90)
    print(name, 'rc', r.returncode, r.stderr[-1500:])
a, b = torch.load('/tmp/cmp_new.pt'), torch.load('/tmp/cmp_old.pt')
for k in a:
    d = (a[k][1] - b[k][1]).norm() / b[k][1].norm()
    q = (a[k][1] - b[k][1]).view(k[0], k[1], 3, -1)
    per = [float(q[:, :, i].norm() / b[k][1].view(k[0], k[1], 3, -1)[:, :, i].norm()) for i in range(3)]
    lone = float(q[:, -1].norm() / b[k][1].view(k[0], k[1], 3, -1)[:, -1].norm())
    print(k, 'dqkv rel_l2 %.3e (dq %.2e dk %.2e dv %.2e, last token %.2e) max_abs %.3e nan %d' % (d, *per, lone, (a[k][1] - b[k][1]).abs().max(), int(torch.isnan(a[k][1]).sum())))
PY
cat gpurun_out/r2_${tag}_cmp.txt
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_block_ops_gpu.py -q -x -k "favor or relu or linear or fused or block" 2>&1 | tail -8 > gpurun_out/r2_${tag}_tests.txt
cat gpurun_out/r2_${tag}_tests.txt
timeout 120 python tools/time_la.py 2>&1 | tee gpurun_out/r2_${tag}_time.txt
