#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=${1:-r2s}
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "cost or sinkhorn or assign or ot or balanc" > $O/${T}_ot_tests.log 2>&1
tail -3 $O/${T}_ot_tests.log
timeout 900 python bench.py --config c5 > $O/${T}_c5.log 2> $O/${T}_c5.err
tail -c 300 $O/${T}_c5.err
python - $O/${T}_c5.log <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
for e in d['sweep']:
    c=e.get('cpu_float64_sinkhorn',{})
    print(f"n={e['n']:>9} k={e['k']:>3} us/iter {e['us_per_iter']:8.1f} frac {e['hbm_frac_per_gpu']:.2f} | cost {e['cost_matrix_ms']*1e3:7.1f}us frac {e['cost_hbm_frac']:.2f} | cpu {c.get('us_per_iter',0)/1e3:.1f} ms/iter {c.get('status')}")
PY
