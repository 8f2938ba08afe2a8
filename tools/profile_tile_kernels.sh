#!/bin/bash
# ncu --set full of the softmax / KERPLE tile kernels at BASELINE configs 1 (B=4096) and 3 (B=1024)
mkdir -p gpurun_out
python bench.py --workload config1 --batch 4096 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/tile_plain1.log 2>&1 || exit 1
python bench.py --workload config3 --batch 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/tile_plain3.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tile_bwd_dkv_kernel -s 3 -c 1 -f -o gpurun_out/prof_tile_softmax_dkv python bench.py --workload config1 --batch 4096 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tile1.log 2>&1

ncu --set full --clock-control none --import-source on -k regex:tile_bwd_dkv_kernel -s 3 -c 1 -f -o gpurun_out/prof_tile_kerple_dkv python bench.py --workload config3 --batch 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tile3.log 2>&1
ls -la gpurun_out | grep prof_tile
