#!/bin/bash
# ncu --set full of the attention kernels of the default bench (after the same command exited 0 without ncu)
TAG=${1:-r2}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
for spec in "la_tc2_fwd_kernel fwd" "la_tc2_bwd_kernel bwd"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$1 -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_$2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_${TAG}_$2.log 2>&1
done
ls -la gpurun_out | grep prof_$TAG
