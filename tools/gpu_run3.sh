#!/bin/bash
# multi-GPU session: decomposed tests across devices, torchrun bench, one-process driver
set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_decomposed.py -m gpu -q -x > gpurun_out/gputests_mg_$N.log 2>&1; tail -5 gpurun_out/gputests_mg_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 120 --warmup 12 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; tail -c 2500 gpurun_out/bench_$N.json; tail -5 gpurun_out/bench_$N.err
timeout 300 python tools/mg_check.py rbc $N 8 > gpurun_out/mg_${N}gpu.log 2>&1; tail -12 gpurun_out/mg_${N}gpu.log
