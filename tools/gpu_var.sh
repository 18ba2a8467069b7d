#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for v in base "$@"; do
  if [ $v = base ]; then unset URE_LIB; else export URE_LIB=$PWD/ultrare_b200/csrc/_obj/var_$v.so; fi
  timeout 300 python tools/prof_mf.py --mode owner --epochs 50 --reps 4 > $O/var_$v.log 2>&1
  echo "== $v"; grep -E "rep 3|losses|Error|error" $O/var_$v.log | cut -c1-250
done
