#!/bin/bash
# round-2 GPU pass F
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > $O/r2f_gpu_tests.log 2>&1
tail -15 $O/r2f_gpu_tests.log
for cfg in "1000000 8" "1000000 32" "10000000 8" "1000000 128" "10000000 32"; do
  set -- $cfg
  timeout 300 python tools/prof_ot.py --n $1 --k $2 --d 64 > $O/r2f_ot_$1_$2.log 2>&1
  URE_COST_WS=0 timeout 300 python tools/prof_ot.py --n $1 --k $2 --d 64 > $O/r2f_ot_$1_$2_nows.log 2>&1
  tail -1 $O/r2f_ot_$1_$2.log | cut -c1-300; tail -1 $O/r2f_ot_$1_$2_nows.log | cut -c1-200
done
timeout 900 python bench.py --config c5 > $O/r2f_c5_1.log 2> $O/r2f_c5_1.err
tail -c 300 $O/r2f_c5_1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cost_ws_kernel|colsum_kernel" -c 4 -o $O/r2f_ot -f \
  python tools/prof_ot.py --n 1000000 --k 8 --d 64 --reps 1 --iters 2 > $O/r2f_ncu_ot.log 2>&1
ls -la $O | tail -6
