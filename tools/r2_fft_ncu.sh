#!/bin/bash
# ncu --set full of the KERPLE FFT kernel (after the plain run exited 0)
python tools/kfft_run.py 8 256 > gpurun_out/kfft_plain.log 2>&1 || { cat gpurun_out/kfft_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:kfft_fwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_kfft python tools/kfft_run.py 8 256 > gpurun_out/ncu_r02_kfft.log 2>&1
ls -la gpurun_out/prof_r02_kfft.ncu-rep
