#!/bin/bash
# N-GPU session: short torchrun bench with the e2e phase trace
set -x
mkdir -p gpurun_out
N=${1:-2}; K=${2:-48}
ORBC_BENCH_TRACE=1 timeout ${T:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $K --warmup 6 --no-cpu-baseline > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; tail -c 1500 gpurun_out/bench_$N.json; grep "bench rank" gpurun_out/bench_$N.err
