#!/bin/bash
# phase trace of the pipelined forward (rebuilds the one file with -DERV_TRACE on the GPU box)
set -e
cd efficient-rpe-vit_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda -Xcompiler -fPIC -DERV_TRACE -c erv_linattn_pipe.cu -o build/erv_linattn_pipe.o
nvcc -shared -o ../erv_b200/lib/liberv_b200.so build/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
cd ../..
timeout 120 python tools/trace_pipe.py > gpurun_out/r2_trace_pipe.txt 2>&1 || true
cat gpurun_out/r2_trace_pipe.txt
