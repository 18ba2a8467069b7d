#!/bin/bash
# ncu --set full of ONE kernel of the bench step: tools/gpu_ncu_k.sh <tag> <kernel regex> [launch-skip]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=$1; K=$2; S=${3:-2}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$K -s $S -c 1 -f -o $O/$T \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/${T}_ncu.log 2>&1
tail -3 $O/${T}_ncu.log
