#!/bin/bash
# hit lists over longer windows, with and without the volume constraint: gpurun -- 'bash tools/gpu_gate2.sh <tag>'
tag=${1:-g2}; out=gpurun_out
for steps in 20 60 240; do for v in "" "--cv"; do n=$(echo "$v" | tr -d ' =-')
  timeout 600 python bench.py --steps $steps --warmup 5 --no-cpu-baseline $v > $out/${tag}_b${steps}_$n.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("$out/${tag}_b${steps}_$n.json").read().strip().splitlines()[-1])
print("$steps", "$v".ljust(6), d["ms_per_step"], d["value"], d["config"].get("hit_lists"))
PY
done; done
