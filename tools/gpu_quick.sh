#!/bin/bash
# quick C2 pass: the bench line with debug timing (+ optional test selection: $2)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-q}
if [ -n "$2" ]; then timeout 900 python -m pytest tests -q -m gpu -x -k "$2" > $O/${T}_tests.log 2>&1; tail -3 $O/${T}_tests.log; fi
URE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-extra > $O/${T}_bench.log 2> $O/${T}_bench.err
tail -c 2200 $O/${T}_bench.err
python - $O/${T}_bench.log <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("ms_per_step", d['ms_per_step'], "e2e", d['e2e']['ms_per_step'], "kernel", d['roofline']['kernel_ms'], "whole", d['roofline']['whole_step_frac'], d['roofline']['whole_step_frac_e2e'])
PY
