#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; T=r2m1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_surface.py -q -m gpu -k "owner or mf or sisa or scratch or instance" -x > $O/${T}_tests.log 2>&1
tail -3 $O/${T}_tests.log
for nt in 4; do
  URE_SCHED_NT=$nt timeout 300 python tools/prof_timeline.py 50 > $O/${T}_tl_nt$nt.log 2>&1
  echo "NT=$nt"; grep -E "owner_schedule|event-timed" $O/${T}_tl_nt$nt.log | cut -c1-200
done
URE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > $O/${T}_bench.log 2> $O/${T}_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2m1_bench.log') if l.startswith('{')][-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['whole_step_frac'])
PY
