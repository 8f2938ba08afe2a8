#!/bin/bash
# Regenerates the round-2 summaries under profiles/ from the artefacts of `tools/r2_evidence.sh <tag>` + `tools/r2_fft_ncu.sh`
# + `tools/time_kerple.py` in gpurun_out/ (run in the build container; ncu -i reads the .ncu-rep files here).
TAG=${1:-r02f}
G=gpurun_out
HEAD=$(git rev-parse --short HEAD)
rep() { local f=$G/prof_${TAG}_$1.ncu-rep; [ -f $f ] || f=$G/prof_r02j_$1.ncu-rep; [ -f $f ] || f=$G/prof_r02i_$1.ncu-rep; echo $f; }  # kernels unchanged since r02i were not re-captured
python tools/ncu_summary.py $(rep fwd) $(rep bwd) $(rep l4bwd) > /tmp/sum_lin.md
python tools/ncu_summary.py $(rep k3fwd) $(rep k3bwd) $(rep k5fwd) $(rep k5bwd) > /tmp/sum_k.md
python tools/ncu_summary.py $(rep s4fwd) $(rep s4bwd) > /tmp/sum_s.md
python tools/ncu_summary.py $G/prof_r02_kfft.ncu-rep > /tmp/sum_fft.md
{
echo "# Round 2: ncu --set full of the linear-attention kernels (B200, clocks free, evidence set $TAG)"
echo
echo "Command per capture (after the same command exited 0 without ncu): \`ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs [--workload config4 --batch 256]\` (tools/r2_evidence.sh $TAG)."
echo "Config 2 (B=1024, N=65, M=256, fp32): \`la_pipe_fwd_kernel\`, \`la_pipe_bwd_kernel\`; config 4 (B=256, N=197, bf16): \`la_tc_bwd_kernel\`."
echo
cat /tmp/sum_lin.md
echo
echo "Round 1 for comparison (profiles/r01_final_ncu_attention.md): forward 105 us, tensor pipe 11 %, XU 18 %, issue 34 %; backward 286 us, tensor pipe 6 %, issue 35 %, no_inst the top stall (16 %)."
echo
echo "## Warp-stall samples by reason (source page, all lines)"
for t in fwd bwd; do echo; echo "### la_pipe_${t}_kernel"; echo '```'; python tools/ncu_stalls.py $(rep $t); echo '```'; done
echo
echo "## Top source lines by stall samples"
for t in fwd bwd; do echo; echo "### la_pipe_${t}_kernel"; echo '```'; python tools/ncu_lines.py $(rep $t) 14; echo '```'; done
echo
echo "Reading: the tensor pipe went from 6 % to 17 % active in the backward and from 11 % to 19 % in the forward; the issue slots are 46-52 % busy."
echo "The largest single stall site of the backward is the MMA-completion wait of the compute warps (\`mbar_try_wait\`, ~14 % of samples, counted as long_scoreboard);"
echo "the polling loops are 15-20 % of the executed warp instructions (a suspend-time hint on try_wait was measured and made the step slower, erv_umma.cuh)."
echo "The instruction-cache stall (no_inst) fell from 16 % to 7 %.  DRAM traffic of the backward is 88 MB against 68 MB algorithmic (the saved [S|z] state)."
} > profiles/r02_ncu_attention.md
{
echo "# Round 2: ncu --set full of the softmax / KERPLE tile kernels on tcgen05 (B200, evidence set $TAG)"
echo
echo "Same recipe as profiles/r02_ncu_attention.md (tools/r2_evidence.sh $TAG).  Columns: config 3 (KERPLE, N=65, M=44, B=1024) forward / backward, config 5 (KERPLE, N=4097, M=44, B=2) forward / backward."
echo
cat /tmp/sum_k.md
echo
echo "Config 4b (softmax + RoPE, N=197, bf16, B=256) forward / backward:"
echo
cat /tmp/sum_s.md
echo
echo "SASS check (cuobjdump -sass of liberv_b200.so): ktile_* and stile_* contain UTCHMMA + LDTM (tcgen05.mma / tcgen05.ld): ktile_fwd 47 / 5, ktile_bwd 57-71 / 5-7, stile_fwd 37 / 5, stile_bwd 27-41 / 3-5; round 1's tile kernels had none."
} > profiles/r02_ncu_tile_kernels.md
{
echo "# Round 2: KERPLE forward, FFT route (erv_kerple_fft.cu) against the Toeplitz-masked tile route (erv_ktile_tc.cu)"
echo
echo "VERDICT round 1 item 4 / north_star kernel 3.  Both routes are built and parity-green (tests/test_parity_gpu.py: test_kerple_fft_route_*);"
echo "the route is chosen per shape from the measurements below (\`kerple_fft_eligible\`: N - 1 > 2048 and (M > 64 or B*H >= 16))."
echo
echo "## Timing (tools/time_kerple.py, CUDA events around the whole forward / backward call, 10 calls after 3 warm-ups, B200)"
echo
echo "The forward time includes every kernel of the route (tile route: W^T / exp tables + the tile kernel; FFT route: tables, the two"
echo "feature-map launches, the feature-pair-major transpose, the coefficient FFT, kfft_fwd_kernel, finalize + CLS kernels).  The backward"
echo "is the tile route in both blocks (it consumes the forward's saved output / normaliser)."
echo
echo '```'
cat $G/${TAG}_time_kerple.txt
echo '```'
echo
echo "## Model"
echo
echo "Per (batch, head) pair the FFT route runs M (Dh + 1) / 2 complex column pairs (two real columns per transform), each one forward and one"
echo "inverse 8192-point transform: 748 transforms at M = 44, 4352 at M = 256, independent of N up to 4097.  One transform is"
echo "3 x 16-point DFTs + 2 twiddle passes per thread (512 threads x 16 points) and two shared-memory exchanges of 64 KB: ~1000 thread"
echo "instructions, ~4.2 k cycles measured (7.9 k before the phi rows were staged as bulk copies from feature-pair-major rows and the filter kept in shared memory)."
echo "At 148 SMs that is 748 x 4.2 k / 148 = 21 k cycles = 10.8 us per pair at M = 44 when the grid is full; measured 22.7 us at 64 pairs including"
echo "the feature maps and the finalize kernels.  The tile route does 2 N^2 (M + Dh) FLOP per pair on the tensor pipe (2.05 GFLOP at N = 4097,"
echo "M = 44: 3.7 us at the bf16 peak with the three-term split) but spends ~5 us per 128 x 128 tile with its phases serialised: 35.8 us per pair."
echo "For M > 64 the tile route has no tcgen05 instance and runs on CUDA cores (510 us per pair at M = 256)."
echo
echo "## ncu --set full of kfft_fwd_kernel<float> (B = 8, N = 4097, M = 256: 816 CTAs of 512 threads; tools/r2_fft_ncu.sh)"
echo
cat /tmp/sum_fft.md
echo
echo '```'
python tools/ncu_stalls.py $G/prof_r02_kfft.ncu-rep
echo '```'
echo
echo "Achieved HBM throughput is ~1 % of peak and the tensor pipe is idle: the kernel is bound by fp32 instruction issue (issue slots 70 % busy, FMA pipe 58 %,"
echo "not_selected the top stall reason, one 16-warp CTA per SM at 128 registers); phi rows are staged from L2 by the TMA unit (cp.async.bulk + mbarrier, UBLKCP in SASS), the filter coefficients stay in shared memory (long_scoreboard 38 % -> 1.8 %)."
} > profiles/r02_kerple_fft_vs_tile.md
python tools/ll_summary.py $G/launches_$TAG.csv 30 > /tmp/ll.txt
{
echo "# Round 2: launch list of the default bench step (config 2, B=1024), evidence set $TAG"
echo
echo "\`ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs\` (after the same command exited 0 without ncu)."
echo "Per-launch times under ncu are cold-cache and serialised; the SHARE of the step is what carries over.  \`spin_kernel\` is \`torch.cuda._sleep\` of the roofline pass (not part of a step)."
echo
echo '```'
cat /tmp/ll.txt
echo '```'
python - "$G/launches_$TAG.csv" <<'PY'
import csv, collections, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
agg=collections.defaultdict(float); tot=0
for row in csv.DictReader(lines):
    try: v=float(row['Metric Value'].replace(',',''))
    except Exception: continue
    k=row['Kernel Name']
    if 'spin_kernel' in k: continue
    tot+=v
    g='attention backward' if 'la_pipe_bwd' in k else 'attention forward' if 'la_pipe_fwd' in k else 'block kernels (ln_qkv, mlp)' if ('mlp_' in k or 'ln_qkv' in k) else 'partial-sum reductions' if 'sum_partials' in k else 'embedding' if 'embed' in k else 'head + loss' if 'head_loss' in k else 'adam' if 'adam' in k else 'library (adds, fills, rng)'
    agg[g]+=v
print("\nShare of the step (spin kernel excluded):\n")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1]): print("* %s: %.1f %%" % (k, 100*v/tot))
PY
} > profiles/r02_launches.md
python - "$G/bench_$TAG.json" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
json.dump(d, open('profiles/r02_bench.json','w'), indent=1)
print('bench', d['value'], d['ms_per_step'])
PY
cp $G/${TAG}_time_kerple.txt profiles/r02_kerple_routes_timing.txt
echo "profiles regenerated from $TAG at $HEAD"
