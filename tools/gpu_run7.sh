#!/bin/bash
# 1-GPU session: the whole GPU suite, then a short bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; tail -30 gpurun_out/gputests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
