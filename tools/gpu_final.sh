#!/bin/bash
# Final single-GPU session of a round: bench line as the driver runs it, launch list under ncu, full captures of the pair kernels.
tag=${1:-r02}; out=gpurun_out; mkdir -p $out
python tools/make_states.py rbc --opt 100 > $out/${tag}_state.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.json
timeout 900 python bench.py --steps 240 --warmup 24 --no-cpu-baseline > $out/${tag}_bench_240.json 2>> $out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err; tail -c 400 $out/${tag}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
# the pair kernels one by one: search (mode 2), record (mode 1), walk (the gate's own choice on repeated evaluations).  pair_only.py runs
# 4 MD steps first (6 launches of k_pair_ll_r, 2 of k_pair_ll_list, 8 of k_pair_prot*), then 3 force evaluations in the forced mode,
# each of which launches the selected kernel and one that returns at once: skip the warm-up launches, capture one evaluation
cap() { # name, kernel regex, debug_nl_mode, launches to skip, launches to capture
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c $5 -o $out/${tag}_prof_$1 -f python tools/pair_only.py rbc 3 2 1 $3 > $out/${tag}_ncu_$1.log 2>&1; tail -1 $out/${tag}_ncu_$1.log
}
cap search k_pair_ll_r 2 6 2
cap record k_pair_ll_r 1 6 2
cap walk k_pair_ll_list -1 2 1
cap prot k_pair_prot 2 8 2
timeout 300 python tools/kernel_bench.py rbc 10 > $out/${tag}_kernel_bench.txt 2>&1
timeout 300 python tools/ll_bench.py rbc 10 > $out/${tag}_ll_bench.txt 2>&1
timeout 300 python tools/gate_cost.py rbc > $out/${tag}_gate_cost.txt 2>&1
timeout 300 python tools/init_bench.py > $out/${tag}_init_bench.txt 2>&1
