#!/bin/bash
# final ncu captures of the round (each after its plain command has run without ncu)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $O/r2z_plain_bench.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2z_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $O/r2z_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mf_owner_kernel|owner_schedule_tab" -c 2 -o $O/r2z_owner -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $O/r2z_ncu_owner.log 2>&1
timeout 300 python tools/prof_ot.py --n 10000000 --k 8 --d 64 --reps 1 --iters 2 > $O/r2z_plain_ot.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cost_ws_kernel" -c 1 -o $O/r2z_cost -f \
  python tools/prof_ot.py --n 10000000 --k 8 --d 64 --reps 1 --iters 2 > $O/r2z_ncu_cost.log 2>&1
ls -la $O/r2z_*
