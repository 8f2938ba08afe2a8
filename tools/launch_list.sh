#!/bin/bash
# short ncu launch list of the default bench (eager warm-up steps + graph replays): gpurun_out/launches_$1.csv
TAG=${1:-x}
ncu --metrics gpu__time_duration.sum --clock-control none -c ${2:-1500} --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll_$TAG.log 2>&1
