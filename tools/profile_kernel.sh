#!/bin/bash
# ncu --set full capture of one kernel of the default bench: tools/profile_kernel.sh <kernel-regex> <tag> [skip]
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-4} -c 1 -f -o gpurun_out/prof_$2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$2.log 2>&1
ls -la gpurun_out/prof_$2.ncu-rep
