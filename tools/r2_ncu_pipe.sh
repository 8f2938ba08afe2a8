#!/bin/bash
# ncu --set full of the pipelined forward (driver: tools/time_la.py, which must exit 0 without ncu first)
TAG=${1:-pipe}
python tools/time_la.py > gpurun_out/time_$TAG.txt 2>&1 || { tail -5 gpurun_out/time_$TAG.txt; exit 1; }
cat gpurun_out/time_$TAG.txt
ncu --set full --clock-control none --import-source on -k regex:la_pipe_fwd_kernel -s 3 -c 1 -f -o gpurun_out/prof_$TAG python tools/time_la.py > gpurun_out/ncu_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
