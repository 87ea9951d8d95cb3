#!/bin/bash
# One GPU session: parity tests, kernel timings, launch list and one full ncu capture of the lipid pair kernel.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'      (outputs under gpurun_out/)
tag=${1:-run}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_host_driver.py 2>&1 | tail -30 > $out/${tag}_tests.log
tail -5 $out/${tag}_tests.log
timeout 600 python tools/ll_bench.py rbc 10 > $out/${tag}_ll_bench.txt 2>&1; cat $out/${tag}_ll_bench.txt | tail -8
timeout 600 python tools/kernel_bench.py rbc 10 > $out/${tag}_kernel_bench.txt 2>&1; tail -45 $out/${tag}_kernel_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair_ll_t -s 2 -c 1 -o $out/${tag}_prof_tile -f python tools/pair_only.py rbc 3 > $out/${tag}_ncu.log 2>&1; tail -3 $out/${tag}_ncu.log
