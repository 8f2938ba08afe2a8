#!/bin/bash
# Round-2 evidence for profiles/: plain bench (all configs), ncu --set full of the attention kernels of configs 2, 3, 4b, 5,
# launch list of the default bench.  Every ncu pass runs after the same command exited 0 without ncu.
TAG=${1:-r02a}
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { tail -5 gpurun_out/bench_$TAG.err; exit 1; }
tail -1 gpurun_out/bench_$TAG.json | cut -c1-300
cap() {  # kernel-regex tag skip bench-args...   (tags listed in $ERV_SKIP_CAPS are not captured: gpurun returns at most 64 MiB)
  local k=$1 t=$2 s=$3; shift 3
  case " $ERV_SKIP_CAPS " in *" $t "*) return;; esac
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs "$@" > gpurun_out/plain_${TAG}_$t.log 2>&1 || { echo "plain $t failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/prof_${TAG}_$t python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs "$@" > gpurun_out/ncu_${TAG}_$t.log 2>&1
}
cap la_pipe_fwd_kernel fwd 4
cap la_pipe_bwd_kernel bwd 4
cap ktile_fwd_kernel k3fwd 3 --workload config3 --batch 1024
cap ktile_bwd_kernel k3bwd 3 --workload config3 --batch 1024
cap stile_fwd_kernel s4fwd 3 --workload config4b --batch 256
cap stile_bwd_kernel s4bwd 3 --workload config4b --batch 256
cap ktile_fwd_kernel k5fwd 3 --workload config5 --batch 2
cap ktile_bwd_kernel k5bwd 3 --workload config5 --batch 2
cap la_tc_bwd_kernel l4bwd 3 --workload config4 --batch 256
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_ll_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
