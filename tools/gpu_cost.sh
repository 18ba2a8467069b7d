#!/bin/bash
# cost-kernel pass: parity tests of the OT kernels, then the C5 points
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-r2t}
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "cost or sinkhorn or assign or ot" > $O/${T}_ot_tests.log 2>&1
tail -4 $O/${T}_ot_tests.log
for cfg in "1000000 8 64" "1000000 32 64" "1000000 64 64" "1000000 128 64" "10000000 8 64" "10000000 128 64" "1000000 128 32" "1000000 64 128"; do
  set -- $cfg
  timeout 300 python tools/prof_ot.py --n $1 --k $2 --d $3 > $O/${T}_ot_$1_$2_$3.log 2>&1
  tail -1 $O/${T}_ot_$1_$2_$3.log | cut -c1-330
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/${T}_launches.csv python tools/prof_ot.py --n 1000000 --k 8 --d 64 --reps 1 --iters 5 > /dev/null 2>&1
cut -d, -f5,15 $O/${T}_launches.csv | tail -30 | cut -c1-120
