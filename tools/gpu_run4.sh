#!/bin/bash
# 1-GPU session: tests, per-kernel timings incl. variants and owned-fraction scaling, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; tail -15 gpurun_out/gputests.log
timeout 300 python tools/kernel_bench.py rbc 10 > gpurun_out/kernel_bench.log 2>&1; cat gpurun_out/kernel_bench.log
timeout 300 python bench.py --steps 120 --warmup 12 --no-cpu-baseline > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err; tail -c 1500 gpurun_out/bench_1.json
