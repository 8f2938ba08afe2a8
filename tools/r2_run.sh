#!/bin/bash
# round-2 GPU check: tests (all failures listed), golden diagnostics, short bench.  Output under gpurun_out/r2_<tag>_*.
tag=${1:-x}
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_${tag}_tests.txt
python tools/diag_golden.py attn_d_relu_circulant_string.npz attn_d_relu_none.npz attn_d_favor_plus_circulant_string.npz > gpurun_out/r2_${tag}_diag.txt 2>&1
ERV_DISABLE_TC2=1 python tools/diag_golden.py attn_d_relu_circulant_string.npz attn_d_relu_none.npz >> gpurun_out/r2_${tag}_diag.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_${tag}_bench.json 2> gpurun_out/r2_${tag}_bench.err
tail -12 gpurun_out/r2_${tag}_tests.txt
cat gpurun_out/r2_${tag}_diag.txt
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_${tag}_bench.json'))
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['per_call_ms'])
    for k,v in d['other_configs'].items(): print(k, {kk:v.get(kk) for kk in ('value','ms_per_step','error')}, (v.get('roofline') or {}).get('per_call_ms'))
    print('cpu',d['cpu_baseline'])
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/r2_${tag}_bench.err').read()[-2000:])
PY
