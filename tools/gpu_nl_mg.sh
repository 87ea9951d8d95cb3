#!/bin/bash
# hit lists on a decomposed run, on / off: gpurun --gpus N -- 'bash tools/gpu_nl_mg.sh <tag> <N> [steps]'
tag=${1:-nlmg}; n=${2:-2}; steps=${3:-60}; out=gpurun_out; mkdir -p $out
python tools/make_states.py rbc --opt 100 > /dev/null 2>&1
for v in 0 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps $steps --warmup 5 --no-cpu-baseline --opt nl_reuse=$v > $out/${tag}_${n}gpu_nl$v.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("$out/${tag}_${n}gpu_nl$v.json").read().strip().splitlines()[-1])
print("N=$n nl_reuse=$v", d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["hit_lists"])
PY
done
