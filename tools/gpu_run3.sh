#!/bin/bash
# multi-GPU session: decomposed tests, one-process decomposed run with kernel report, torchrun bench
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_decomposed.py -m gpu -q -x > gpurun_out/gputests_$N.log 2>&1; tail -5 gpurun_out/gputests_$N.log
timeout 300 python tools/mg_check.py rbc $N 8 > gpurun_out/mg_${N}gpu.log 2>&1; head -16 gpurun_out/mg_${N}gpu.log; grep -A24 "^rank 0:" gpurun_out/mg_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 240 --warmup 24 --no-cpu-baseline > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; tail -c 2500 gpurun_out/bench_$N.json; tail -5 gpurun_out/bench_$N.err
