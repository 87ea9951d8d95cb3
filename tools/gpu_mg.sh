#!/bin/bash
# Multi-GPU session: gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_mg.sh <tag> <N>'
tag=${1:-mg}; n=${2:-2}
out=gpurun_out; mkdir -p $out
python tools/make_states.py rbc --opt 100 > $out/${tag}_state.log 2>&1
ORBC_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
tail -c 1500 $out/${tag}_bench_${n}gpu.json; grep -E "e2e|upload|phases" $out/${tag}_bench_${n}gpu.err | tail -6
timeout 600 python tools/mg_check.py rbc $n 8 > $out/${tag}_mg_check_${n}.txt 2>&1; tail -45 $out/${tag}_mg_check_${n}.txt
