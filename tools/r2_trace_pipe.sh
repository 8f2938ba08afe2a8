#!/bin/bash
# phase trace of the pipelined forward / backward (rebuilds the two files with -DERV_TRACE on the GPU box)
set -e
cd efficient-rpe-vit_b200/csrc
for f in erv_linattn_pipe erv_linattn_pipe_bwd; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda -Xcompiler -fPIC -DERV_TRACE -c $f.cu -o build/$f.o
done
nvcc -shared -o ../erv_b200/lib/liberv_b200.so build/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
cd ../..
for w in ${1:-fwd bwd}; do
  timeout 120 python tools/trace_pipe.py $w > gpurun_out/r2_trace_pipe_$w.txt 2>&1 || true
  cat gpurun_out/r2_trace_pipe_$w.txt
done
