#!/bin/bash
# round-2 baseline probe: host profile of the C2 step, c4small LAZY throughput + ncu of the lazy kernel
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi -L > $O/r2_smi.txt 2>&1
timeout 300 python tools/prof_host.py 50 > $O/r2_prof_host.log 2>&1
URE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > $O/r2_bench0.log 2> $O/r2_bench0.err
timeout 600 python tools/run_config.py --config c4small --mode lazy --max-steps 300 > $O/r2_c4small_lazy.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mf_train_lazy_kernel -c 1 -o $O/r2_lazy_c4small -f \
  python tools/run_config.py --config c4small --mode lazy --max-steps 40 > $O/r2_ncu_lazy.log 2>&1
ls -la $O | tail -20
