#!/bin/bash
# compute-sanitizer over the hot path on a small fixture (memcheck single GPU + 2 ranks, racecheck single GPU).
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh <tag>'
tag=${1:-san}; out=gpurun_out; mkdir -p $out
S=/usr/local/cuda/bin/compute-sanitizer
ndev=$(nvidia-smi -L | wc -l); [ "$ndev" -gt 2 ] && ndev=2      # 2 ranks: one device each when the box has two
timeout 600 $S --tool memcheck --error-exitcode 9 python tools/sanitize_driver.py 1 6 > $out/${tag}_memcheck_single.txt 2>&1; echo "memcheck single rc=$?"; tail -4 $out/${tag}_memcheck_single.txt
timeout 600 $S --tool memcheck --error-exitcode 9 python tools/sanitize_driver.py 2 6 ndev=$ndev > $out/${tag}_memcheck_2ranks.txt 2>&1; echo "memcheck 2 ranks rc=$?"; tail -4 $out/${tag}_memcheck_2ranks.txt
timeout 900 $S --tool racecheck --error-exitcode 9 python tools/sanitize_driver.py 1 4 > $out/${tag}_racecheck_single.txt 2>&1; echo "racecheck single rc=$?"; tail -4 $out/${tag}_racecheck_single.txt
grep -E "Invalid|Race|ERROR SUMMARY|at .*\(" $out/${tag}_memcheck_2ranks.txt | head -20
