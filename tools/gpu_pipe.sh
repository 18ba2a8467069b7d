#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-r2pp}
timeout 600 python -m pytest tests/test_gpu_surface.py -q -m gpu -x -k "optimistic or full_size or instance" > $O/${T}_tests.log 2>&1
tail -12 $O/${T}_tests.log
for g in 1 2 3 5; do
  URE_PIPE_GROUPS=$g timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-extra > $O/${T}_bench.log 2> $O/${T}_bench.err
  python - $O/${T}_bench.log $g <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("groups", sys.argv[2], "ms_per_step %.4f e2e %.4f kernel %.4f" % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms']))
PY
done
URE_PIPE_GROUPS=2 timeout 300 python tools/prof_timeline.py 50 e2e 2>&1 | grep -A30 start_us | cut -c1-125
