#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-r2v}
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "sinkhorn or ot" > $O/${T}_ot_tests.log 2>&1
tail -3 $O/${T}_ot_tests.log
for p in 0 1; do
for cfg in "1000000 32 64" "2000000 8 64" "10000000 8 64" "10000000 32 64" "1000000 128 64"; do
  set -- $cfg
  URE_SK_PERSIST=$p timeout 300 python tools/prof_ot.py --n $1 --k $2 --d $3 > $O/${T}_p${p}_$1_$2.log 2>&1
  python - $O/${T}_p${p}_$1_$2.log $p <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("persist", sys.argv[2], d['n'], d['k'], "us/iter %.1f frac %.2f" % (d['sinkhorn_persistent']['us_per_iter'], d['sinkhorn_persistent']['GBps']/6550.4))
PY
done; done
