#!/bin/bash
# compute-sanitizer over the tcgen05 / mbarrier kernels: the UMMA probe tests and the golden parity cases that dispatch to the
# pipelined short-sequence kernels (shape d), the long-sequence kernels and the softmax / KERPLE tile kernels (shape e)
SEL='tests/test_umma_gpu.py tests/test_parity_gpu.py::test_attention_fp32_matches_reference'
for tool in memcheck synccheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest $SEL -q -x -k "attn_d or attn_e or (probe and 128-128)" > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.txt | tail -4
done
