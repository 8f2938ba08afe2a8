#!/bin/bash
# phase trace of the short-sequence backward (rebuilds the one file with -DERV_TRACE on the GPU box, then restores it)
set -e
cd efficient-rpe-vit_b200/csrc
cp build/erv_linattn_tc2_bwd.o /tmp/bwd_keep.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda -Xcompiler -fPIC -DERV_TRACE -c erv_linattn_tc2_bwd.cu -o build/erv_linattn_tc2_bwd.o
nvcc -shared -o ../erv_b200/lib/liberv_b200.so build/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
cd ../..
python tools/trace_bwd.py > gpurun_out/r2_trace_bwd.txt 2>&1 || true
cp /tmp/bwd_keep.o efficient-rpe-vit_b200/csrc/build/erv_linattn_tc2_bwd.o
tail -60 gpurun_out/r2_trace_bwd.txt
