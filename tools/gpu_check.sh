#!/bin/bash
# One GPU session: parity tests, bench line, kernel timings, microbenchmarks and (optionally) full ncu captures of named kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh <tag> [kernel regex for ncu ...]'      (outputs under gpurun_out/)
tag=${1:-run}; shift
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
[ -x tools/_build/fp32_peak ] && tools/_build/fp32_peak > $out/${tag}_fp32_peak.json 2>&1
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_host_driver.py 2>&1 | grep -v " s on " > $out/${tag}_tests.log
grep -E "^(FAILED|ERROR)|passed|failed" $out/${tag}_tests.log | tail -12
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 3000 $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
timeout 600 python tools/ll_bench.py rbc 10 > $out/${tag}_ll_bench.txt 2>&1
for k in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $out/${tag}_prof_$k -f python tools/pair_only.py rbc 3 2 1 > $out/${tag}_ncu_$k.log 2>&1; tail -1 $out/${tag}_ncu_$k.log
done
