#!/bin/bash
# 1-GPU validation session: GPU tests, smoke, bench (both arms), per-kernel timings
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; tail -4 gpurun_out/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
timeout 300 python tools/kernel_bench.py rbc 10 > gpurun_out/kernel_bench.log 2>&1; tail -40 gpurun_out/kernel_bench.log
