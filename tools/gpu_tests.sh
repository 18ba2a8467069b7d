#!/bin/bash
# round-2 GPU pass B: full GPU test-suite, bench line, OT profile, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=10 > $O/r2b_gpu_tests.log 2>&1
tail -30 $O/r2b_gpu_tests.log
URE_BENCH_DEBUG=1 timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2b_bench1.log 2> $O/r2b_bench1.err
tail -c 2500 $O/r2b_bench1.err
for cfg in "1000000 8" "1000000 32" "10000000 8"; do
  set -- $cfg
  timeout 300 python tools/prof_ot.py --n $1 --k $2 --d 64 > $O/r2b_ot_$1_$2.log 2>&1
  tail -1 $O/r2b_ot_$1_$2.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2b_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $O/r2b_ncu_bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2b_bench_ref.log 2>&1
ls -la $O | tail -8
