#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for v in base "$@"; do
  if [ $v = base ]; then unset URE_LIB; else export URE_LIB=$PWD/ultrare_b200/csrc/_obj/var_$v.so; fi
  timeout 600 python tools/run_config.py --config c4small --mode runs --max-steps 100 > $O/runs_$v.log 2>&1
  echo "== $v"; tail -1 $O/runs_$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms'], d['interactions_per_s']/1e9, d['frac_of_hbm_peak'], d['first_epoch_rmse_rank0'])"
done
