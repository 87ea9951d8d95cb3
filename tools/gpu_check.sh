#!/bin/bash
# One GPU session: parity tests, kernel timings, microbenchmarks and (optionally) one full ncu capture of a lipid pair kernel.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh <tag> [ncu kernel regex] [ll_variant for the capture]'      (outputs under gpurun_out/)
tag=${1:-run}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
[ -x tools/_build/fp32_peak ] && tools/_build/fp32_peak > $out/${tag}_fp32_peak.json 2>&1 && cat $out/${tag}_fp32_peak.json
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_host_driver.py 2>&1 | grep -v " s on " > $out/${tag}_tests.log
grep -E "^(FAILED|ERROR)|passed|failed" $out/${tag}_tests.log | tail -12
timeout 600 python tools/ll_bench.py rbc 10 > $out/${tag}_ll_bench.txt 2>&1; tail -60 $out/${tag}_ll_bench.txt
if [ -n "$2" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o $out/${tag}_prof -f python tools/pair_only.py rbc 3 2 ${3:-0} > $out/${tag}_ncu.log 2>&1; tail -3 $out/${tag}_ncu.log
fi
