#!/bin/bash
# one short bench line per BASELINE config (not the headline; for DESIGN.md's table)
mkdir -p gpurun_out
for spec in "config1 128" "config1 4096" "config3 1024" "config4 256" "config4b 256" "config5 2"; do
  set -- $spec
  timeout 280 python bench.py --workload $1 --batch $2 --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/cfg_$1_$2.err | tail -1 > gpurun_out/cfg_$1_$2.json
  python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/cfg_{sys.argv[1]}_{sys.argv[2]}.json"))
    r=d.get("roofline") or {}
    print(sys.argv[1], "B=",sys.argv[2], "img/s=%.1f"%d["value"], "ms/step=%.3f"%d["ms_per_step"], "e2e=%.1f"%d["e2e"]["value"], r.get("per_call_ms"))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e); print(open(f"gpurun_out/cfg_{sys.argv[1]}_{sys.argv[2]}.err").read()[-600:])
PY
done
