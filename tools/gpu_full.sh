#!/bin/bash
# full GPU pass: every -m gpu test, the C2 bench line (debug timing), the step timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-r2n}
timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 > $O/${T}_gpu_tests.log 2>&1
tail -9 $O/${T}_gpu_tests.log
timeout 300 python tools/prof_timeline.py 50 > $O/${T}_tl.log 2>&1
grep -E "event-timed|span" $O/${T}_tl.log | cut -c1-400
URE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > $O/${T}_bench.log 2> $O/${T}_bench.err
tail -c 1500 $O/${T}_bench.err
python - $O/${T}_bench.log <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("ms_per_step", d['ms_per_step'], "e2e", d['e2e']['ms_per_step'], "kernel", d['roofline']['kernel_ms'], "whole", d['roofline']['whole_step_frac'], d['roofline']['whole_step_frac_e2e'])
PY
