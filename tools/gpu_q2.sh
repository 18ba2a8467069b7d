#!/bin/bash
# selected GPU tests ($2), two quick bench lines, the tail of the CUPTI timeline of one step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-q2}
if [ -n "$2" ]; then timeout 1500 python -m pytest tests -q -m gpu -x -k "$2" > $O/${T}_tests.log 2>&1; tail -3 $O/${T}_tests.log; fi
for rep in 1 2; do
URE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-extra > $O/${T}_bench.log 2> $O/${T}_bench.err
python - $O/${T}_bench.log <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("ms_per_step %.3f e2e %.3f kernel %.3f" % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms']))
PY
done
timeout 600 python tools/prof_timeline.py 50 2>&1 | cut -c1-140 | tail -${3:-22}
