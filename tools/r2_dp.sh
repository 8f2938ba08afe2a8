#!/bin/bash
# N-GPU bench lines: one-shot peer-memory all-reduce (default) vs NCCL (ERV_NCCL_ALLREDUCE=1)
N=${1:-2}
for mode in peer nccl; do
  if [ $mode = nccl ]; then export ERV_NCCL_ALLREDUCE=1; else unset ERV_NCCL_ALLREDUCE; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --no-other-configs > gpurun_out/r2_dp_${mode}_n$N.json 2> gpurun_out/r2_dp_${mode}_n$N.err
  echo "$mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_dp_${mode}_n$N.json').read().strip().splitlines()[-1])
    print('$mode', 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'loss', d['final_loss'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r2_dp_${mode}_n$N.err').read()[-2500:])
PY
done
python bench.py --steps 50 --warmup 5 --no-other-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'], d['roofline']['per_call_ms'])"
