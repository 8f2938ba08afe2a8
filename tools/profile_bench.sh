#!/bin/bash
# ncu evidence for bench.py's default workload (run AFTER the same command exited 0 without ncu):
#   1. launch list (gpu__time_duration per launch)      -> gpurun_out/launches_$TAG.csv
#   2. --set full capture of the attention fwd / bwd     -> gpurun_out/prof_$TAG_{fwd,bwd}.ncu-rep
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
for k in fwd bwd; do
  ncu --set full --clock-control none --import-source on -k regex:la_tc2_${k}_kernel -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_$k $CMD > gpurun_out/ncu_${TAG}_$k.log 2>&1
done
ls -la gpurun_out | grep $TAG
