#!/bin/bash
# Measurements of the BASELINE.json configurations that are not the default bench line (row g1 of the round-1 verdict).
#   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_configs.sh <tag> <N> [what...]'       what: cv weak patch nh
tag=$1; n=$2; shift 2
out=gpurun_out; mkdir -p $out
run() {  # label, bench arguments...
  label=$1; shift
  if [ "$n" -gt 1 ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --no-cpu-baseline "$@" 2>$out/${tag}_${label}_${n}gpu.err | tail -1 > $out/${tag}_${label}_${n}gpu.json
  else
    timeout 900 python bench.py --no-cpu-baseline "$@" 2>$out/${tag}_${label}_${n}gpu.err | tail -1 > $out/${tag}_${label}_${n}gpu.json
  fi
  python - "$out/${tag}_${label}_${n}gpu.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], "value %.4g  %.4f ms/step  e2e %.4g  %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["workload"][:60]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
python tools/make_states.py rbc --opt 100 > $out/${tag}_state.log 2>&1
for what in "$@"; do
  case $what in
    base) run rbc --steps 60 --warmup 12 ;;
    cv) run rbc_cv --steps 60 --warmup 12 --cv ;;
    weak) run weak_patch --steps 40 --warmup 8 --workload patch:$(python -c "print(int(1.05e6 * $n))") --scaling weak ;;
    patch) for s in 1e5 3e5 1e6 3e6 1e7 3e7 5e7; do run patch_$s --steps 40 --warmup 8 --workload patch:$s; done ;;
    nh) timeout 600 python tools/config5_nh_frames.py rbc 400 100 > $out/${tag}_config5_nh_frames.txt 2>&1; tail -4 $out/${tag}_config5_nh_frames.txt ;;
  esac
done
