#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-r2fa}
timeout 900 python -m pytest tests -q -m gpu -x -k "eval_jobs or faithful or scratch or instance or sisa or rank or ensemble" > $O/${T}_tests.log 2>&1
tail -5 $O/${T}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > $O/${T}_bench.log 2> $O/${T}_bench.err
tail -c 300 $O/${T}_bench.err
python - $O/${T}_bench.log <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("ms_per_step", d['ms_per_step'], "e2e", d['e2e']['ms_per_step'], "modes", d['epoch_eval_modes'])
PY
