#!/bin/bash
# full GPU test-suite + optional extra command ($1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 > $O/r2_gpu_tests.log 2>&1
tail -40 $O/r2_gpu_tests.log
if [ -n "$1" ]; then bash -c "$1"; fi
