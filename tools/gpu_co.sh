#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-r2co}
timeout 300 python -m pytest tests/test_gpu_surface.py -q -m gpu -x -k "optimistic or full_size" > $O/${T}_tests.log 2>&1
tail -5 $O/${T}_tests.log
for co in 1 0; do
URE_SCHED_CO=$co URE_BENCH_DEBUG=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > $O/${T}_bench$co.log 2> $O/${T}_bench$co.err
tail -c 300 $O/${T}_bench$co.err
python - $O/${T}_bench$co.log $co <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("CO", sys.argv[2], "ms_per_step", d['ms_per_step'], "e2e", d['e2e']['ms_per_step'], "kernel", d['roofline']['kernel_ms'], "whole", d['roofline']['whole_step_frac'], d['roofline']['whole_step_frac_e2e'])
PY
done
URE_SCHED_CO=1 timeout 300 python tools/prof_timeline.py 50 > $O/${T}_tl.log 2>&1
grep -E "owner_schedule|mf_owner_kernel|event-timed|span" $O/${T}_tl.log | cut -c1-160
