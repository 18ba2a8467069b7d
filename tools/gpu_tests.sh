#!/bin/bash
# round-2 GPU pass E
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > $O/r2e_gpu_tests.log 2>&1
tail -15 $O/r2e_gpu_tests.log
for cfg in "1000000 8" "1000000 32" "10000000 8" "1000000 128"; do
  set -- $cfg
  timeout 300 python tools/prof_ot.py --n $1 --k $2 --d 64 > $O/r2e_ot_$1_$2.log 2>&1
  URE_COST_1ACC=1 timeout 300 python tools/prof_ot.py --n $1 --k $2 --d 64 > $O/r2e_ot_$1_$2_1acc.log 2>&1
  tail -1 $O/r2e_ot_$1_$2.log | cut -c1-300; tail -1 $O/r2e_ot_$1_$2_1acc.log | cut -c1-200
done
URE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > $O/r2e_bench.log 2> $O/r2e_bench.err
tail -c 800 $O/r2e_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cost_tma_kernel" -c 2 -o $O/r2e_cost -f \
  python tools/prof_ot.py --n 1000000 --k 8 --d 64 --reps 1 --iters 2 > $O/r2e_ncu_cost.log 2>&1
ls -la $O | tail -6
