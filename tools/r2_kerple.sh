#!/bin/bash
# KERPLE parity (golden + oracle cases) and per-config timing
tag=${1:-k}
python -m pytest tests/test_parity_gpu.py -q -k "most_general or kerple" 2>&1 | tail -25 > gpurun_out/r2_${tag}_kerple_tests.txt
tail -25 gpurun_out/r2_${tag}_kerple_tests.txt
for cfg in "config3 1024" "config3_p8 1024" "config5 2"; do
  set -- $cfg
  python bench.py --workload $1 --batch $2 --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['roofline']['per_call_ms'])"
done
