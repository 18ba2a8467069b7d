#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for rep in 1 2 3; do
for mode in "URE_SCHED_CO_FIRST=1" "URE_SCHED_CO=0"; do
  env $mode timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > $O/co3_bench.log 2> $O/co3_bench.err
  python - $O/co3_bench.log "$mode" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], "ms_per_step %.4f e2e %.4f kernel %.4f" % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms']))
PY
done; done
