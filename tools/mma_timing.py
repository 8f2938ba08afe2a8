import sys, torch
sys.path.insert(0,'efficient-rpe-vit_b200')
from erv_b200 import _capi as C
lib = C.load()
cyc = torch.zeros(1, dtype=torch.int64, device='cuda')
for bf16 in (0,1):
    for (amn,bmn) in ((0,0),(1,1),(0,1)):
        if not bf16 and (amn or bmn): continue
        for N in (16,32,64,128,256):
            res=[]
            for iters in (1, 8, 64):
                C.check(lib.erv_debug_umma_timing(N, bf16, amn, bmn, iters, C.ptr(cyc), C.stream()))
                C.check(lib.erv_debug_umma_timing(N, bf16, amn, bmn, iters, C.ptr(cyc), C.stream()))
                torch.cuda.synchronize(); res.append(int(cyc.item()))
            per = (res[2]-res[1])/56
            print(f"bf16={bf16} a_mn={amn} b_mn={bmn} N={N:3d}: 1 mma {res[0]:5d} cyc, 8: {res[1]:5d}, 64: {res[2]:6d}  -> {per:6.1f} cyc/mma")
