#!/bin/bash
# Round-end evidence for profiles/: plain bench, ncu --set full of the attention kernels and the MLP backward,
# short launch list, per-config bench lines.  Every ncu pass runs after the same command exited 0 without ncu.
TAG=${1:-r01_final}
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { tail -5 gpurun_out/bench_$TAG.err; exit 1; }
tail -1 gpurun_out/bench_$TAG.json | cut -c1-400
for spec in "la_tc2_fwd_kernel fwd" "la_tc2_bwd_kernel bwd" "mlp_bwd_tc_kernel mlpbwd"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$1 -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_$2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}_$2.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll_$TAG.log 2>&1
bash tools/bench_configs.sh
ncu --set full --clock-control none --import-source on -k regex:tile_bwd_dkv_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_kerple_dkv python bench.py --workload config3 --batch 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}_kerple.log 2>&1
ls -la gpurun_out | grep $TAG
