#!/bin/bash
# GPU session: tests, per-kernel timings, bench, ncu launch list, ncu full capture of the pair kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; tail -3 gpurun_out/gputests.log
python bench.py --steps 120 --warmup 12 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
python tools/kernel_bench.py rbc 10 > gpurun_out/kernel_bench.log 2>&1; cat gpurun_out/kernel_bench.log
python bench.py --impl reference --steps 20 --warmup 2 --ref-budget 60 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 8 -c 2 -f -o gpurun_out/pair_full python tools/pair_only.py rbc 1 > gpurun_out/ncu_pair.log 2>&1
ls -la gpurun_out
