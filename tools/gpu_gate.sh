#!/bin/bash
# hit-list gate check: GPU tests, then the bench at 20 / 60 / 240 steps with and without the lists (and with constrain_volume)
tag=${1:-gate}; out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_host_driver.py 2>&1 | grep -v " s on " > $out/${tag}_tests.log
grep -E "^(FAILED|ERROR)|passed|failed" $out/${tag}_tests.log | tail -12
for steps in 20 60 240; do
  for v in "" "--opt nl_reuse=0" "--cv" "--cv --opt nl_reuse=0"; do
    n=$(echo "$v" | tr -d ' =-')
    timeout 600 python bench.py --steps $steps --warmup 5 --no-cpu-baseline $v > $out/${tag}_b${steps}_$n.json 2> $out/${tag}_b${steps}_$n.err
    python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_b${steps}_$n.json").read().strip().splitlines()[-1])
    print("$steps", "$v".ljust(24), d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"].get("hit_lists"))
except Exception as e: print("$steps $v failed", e)
PY
  done
done
