#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py rbc 10 > gpurun_out/kernel_bench.log 2>&1; tail -32 gpurun_out/kernel_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pair_ll_h|k_cell_bounds" -s 8 -c 2 -f -o gpurun_out/pair_half python tools/pair_only.py rbc 1 > gpurun_out/ncu_pair.log 2>&1; tail -3 gpurun_out/ncu_pair.log
