#!/bin/bash
# 1-GPU session: tests, kernel timings, bench, drop-in program vs reference program
set -x
mkdir -p gpurun_out
ROOT=$(pwd)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; tail -4 gpurun_out/gputests.log
timeout 300 python tools/kernel_bench.py rbc 10 > gpurun_out/kernel_bench.log 2>&1; grep -E "ll_half|ll_variant|rebuild|verlet|run_langevin|impl" gpurun_out/kernel_bench.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1800 gpurun_out/bench.json
# the drop-in program and the reference program, same command line (BASELINE.json configs[1]: the RBC mesh, -E 100 -t 10)
mkdir -p /tmp/run_ours /tmp/run_ref
cd /tmp/run_ours; S=$(date +%s.%N); timeout 600 $ROOT/openrbc_b200/host/_build/openrbc_b200 -i trimesh -m $ROOT/oracle/_ref/example-large/rbc -E 100 -t 10 > $ROOT/gpurun_out/driver_ours.log 2>&1; E=$(date +%s.%N); echo "wall $(echo "$E - $S" | bc) s" >> $ROOT/gpurun_out/driver_ours.log
cd /tmp/run_ref; S=$(date +%s.%N); OMP_PROC_BIND=close OMP_PLACES=cores timeout 900 $ROOT/oracle/_ref/openrbc -i trimesh -m $ROOT/oracle/_ref/example-large/rbc -E 100 -t 10 > $ROOT/gpurun_out/driver_ref.log 2>&1; E=$(date +%s.%N); echo "wall $(echo "$E - $S" | bc) s" >> $ROOT/gpurun_out/driver_ref.log
cd $ROOT
grep -E "steps \*|wall|^[0-9.]+ +\t|optimization|main-loop" gpurun_out/driver_ours.log | tail -16
grep -E "steps \*|wall|^[0-9.]+ +\t|optimization|main-loop" gpurun_out/driver_ref.log | tail -16
