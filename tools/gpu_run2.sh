#!/bin/bash
# GPU session: single-GPU regression + decomposed runs (several ranks on one device)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/gputests.log 2>&1; tail -5 gpurun_out/gputests.log
timeout 600 python -m pytest tests/test_gpu_decomposed.py -m gpu -q > gpurun_out/gputests_mg.log 2>&1; tail -40 gpurun_out/gputests_mg.log
timeout 300 python tools/mg_check.py rbc 2 8 > gpurun_out/mg2.log 2>&1; tail -20 gpurun_out/mg2.log
timeout 300 python tools/mg_check.py rbc 4 8 > gpurun_out/mg4.log 2>&1; tail -20 gpurun_out/mg4.log
