#!/bin/bash
# 1-GPU profiling session: ncu full capture of the pair kernels (one launch each), ncu launch list of a short bench
set -x
mkdir -p gpurun_out
python tools/pair_only.py rbc 0 > /dev/null 2>&1   # generates the state outside the profiler
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_pair_ll_r|k_pair_prot" -c 2 -f -o gpurun_out/pair_v4 python tools/pair_only.py rbc 1 > gpurun_out/ncu_pair.log 2>&1; tail -3 gpurun_out/ncu_pair.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/bench_under_ncu.log
ls -la gpurun_out
