#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_surface.py -q -m gpu -x -k "sinkhorn or balanced or assign or ot_cluster or instance or cost" > $O/r2_ot_tests.log 2>&1
tail -30 $O/r2_ot_tests.log
