"""tcgen05.mma cost with independent accumulators / A in tensor memory / M = 64, plus one-SM unit probes.
Run on the GPU box: python tools/mma_timing2.py > gpurun_out/mma_timing2.txt"""
import sys, torch
sys.path.insert(0, 'efficient-rpe-vit_b200')
from erv_b200 import _capi as C
lib = C.load()
cyc = torch.zeros(2, dtype=torch.int64, device='cuda')
sink = torch.zeros(512, device='cuda')


def t2(N, M, bf16, a_tmem, b_mn, nacc, iters, elected=1):
    for _ in range(2):
        C.check(lib.erv_debug_umma_timing2(N, M, bf16, a_tmem, b_mn, nacc, iters, elected, C.ptr(cyc), C.stream()))
    torch.cuda.synchronize()
    c = cyc.tolist()
    return c[0], c[1]


print("# steady-state SM cycles per tcgen05.mma: (t(72) - t(8)) / 64; issue = issue loop alone")
print("kind M N a_src b_major nacc issue | cyc/mma  issue/mma  single")
for elected in (0, 1):
    for bf16 in (1, 0):
        for M in (128, 64):
            for a_tmem in (0, 1):
                if a_tmem and M == 64:
                    continue
                for b_mn in ((0, 1) if bf16 else (0,)):
                    if b_mn and (M == 64 or not elected):
                        continue
                    for N in (16, 32, 64, 96, 128, 256):
                        for nacc in (1, 2, 4):
                            if nacc * N > 448 or (not elected and nacc > 1):
                                continue
                            a8 = t2(N, M, bf16, a_tmem, b_mn, nacc, 8, elected)
                            a72 = t2(N, M, bf16, a_tmem, b_mn, nacc, 72, elected)
                            a1 = t2(N, M, bf16, a_tmem, b_mn, nacc, 1, elected)
                            print(f"{'bf16' if bf16 else 'tf32'} {M:3d} {N:3d} {'tmem' if a_tmem else 'smem'} "
                                  f"{'mn' if b_mn else 'k '} {nacc} {'elect' if elected else 'tid0 '} | "
                                  f"{(a72[0]-a8[0])/64:7.1f} {(a72[1]-a8[1])/64:7.1f} {a1[0]:6d}")

print("# unit probes, one SM: cycles per instruction per warp-slot (SM-wide instr/clk in brackets)")
names = {0: "tcgen05.ld 32x32b.x32 (128 B/thread)", 1: "tcgen05.st 32x32b.x8 (32 B/thread)", 2: "4 x cvt.rn.bf16x2.f32 (+4 LOP)",
         3: "4 x ex2.approx (+4 FADD)", 4: "4 x ex2 + 4 FADD + 2 cvt", 5: "4 x ffma"}
for mode in range(6):
    for threads in (128, 256, 512):
        res = []
        for iters in (256, 2304):
            for _ in range(2):
                C.check(lib.erv_debug_unit_probe(mode, threads, iters, C.ptr(cyc), C.ptr(sink), C.stream()))
            torch.cuda.synchronize()
            res.append(int(cyc[0].item()))
        per = (res[1] - res[0]) / 2048
        warps = threads // 32
        extra = ""
        if mode == 0:
            extra = f"  -> {threads * 128 / per:7.0f} B/clk/SM"
        if mode == 1:
            extra = f"  -> {threads * 32 / per:7.0f} B/clk/SM"
        print(f"{names[mode]:40s} threads={threads:3d}: {per:7.2f} cyc/iter  [{warps * 32 / per:6.1f} thread-instr/clk]{extra}")
