#!/bin/bash
# A/B of one environment switch on ONE box: tools/gpu_ab.sh <tag> <VAR> <a> <b> [pytest -k selection]
# three alternating bench runs per value (step / e2e ms), then the optional tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-ab}; V=$2; A=$3; B=$4
if [ -n "$5" ]; then timeout 900 python -m pytest tests -q -m gpu -x -k "$5" > $O/${T}_tests.log 2>&1; tail -3 $O/${T}_tests.log; fi
for rep in 1 2 3; do
  for x in "$A" "$B"; do
    env $V=$x timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-extra > $O/${T}_bench.log 2> $O/${T}_bench.err
    python - $O/${T}_bench.log "$V=$x" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], "ms_per_step %.3f e2e %.3f kernel %.3f" % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms']))
PY
  done
done
