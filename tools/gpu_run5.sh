#!/bin/bash
# 1-GPU session: tests, bench (+reference arm), ncu launch list + full capture, drop-in program vs reference program, patch sweep
set -x
mkdir -p gpurun_out
ROOT=$(pwd)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; tail -4 gpurun_out/gputests.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1800 gpurun_out/bench.json
timeout 400 python bench.py --impl reference --steps 40 --warmup 2 --ref-budget 60 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
timeout 300 python tools/kernel_bench.py rbc 10 > gpurun_out/kernel_bench.log 2>&1; grep -E "rebuild|ll_variant|run_langevin|voronoi|cell_update" gpurun_out/kernel_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 8 -c 2 -f -o gpurun_out/pair_full python tools/pair_only.py rbc 1 > gpurun_out/ncu_pair.log 2>&1; tail -2 gpurun_out/ncu_pair.log
# the drop-in program and the reference program, same command line (config 1/2 of BASELINE.json: -E 100 -t 10 on the RBC mesh)
mkdir -p /tmp/run_ours /tmp/run_ref
(cd /tmp/run_ours && /usr/bin/time -v timeout 600 $ROOT/openrbc_b200/host/_build/openrbc_b200 -i trimesh -m $ROOT/oracle/_ref/example-large/rbc -E 100 -t 10 > $ROOT/gpurun_out/driver_ours.log 2>&1)
grep -E "steps \*|Lost|^ *[0-9.]+\s+[0-9.]+\s+[0-9.]+\s+[0-9]+|Elapsed" gpurun_out/driver_ours.log | tail -8
(cd /tmp/run_ref && OMP_PROC_BIND=close OMP_PLACES=cores /usr/bin/time -v timeout 900 $ROOT/oracle/_ref/openrbc -i trimesh -m $ROOT/oracle/_ref/example-large/rbc -E 100 -t 10 > $ROOT/gpurun_out/driver_ref.log 2>&1)
grep -E "steps \*|Lost|^ *[0-9.]+\s+[0-9.]+\s+[0-9.]+\s+[0-9]+|Elapsed" gpurun_out/driver_ref.log | tail -8
for n in 1e5 1e6 1e7 3e7; do
  timeout 600 python bench.py --workload patch:$n --steps 60 --warmup 6 --no-cpu-baseline > gpurun_out/bench_patch_$n.json 2> gpurun_out/bench_patch_$n.err; tail -c 400 gpurun_out/bench_patch_$n.json | head -c 400; echo
done
