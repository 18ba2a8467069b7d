#!/bin/bash
# end-of-round check on N = $1 GPUs: the new single-GPU tests ($2: pytest -k), the NCCL tests, then both bench arms exactly
# as the driver launches them at N GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}; O=gpurun_out
if [ -n "$2" ]; then timeout 600 python -m pytest tests -q -m gpu -x -k "$2" > $O/m2_new_tests.log 2>&1; tail -3 $O/m2_new_tests.log; fi
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $O/m2_multi_tests.log 2>&1; tail -4 $O/m2_multi_tests.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
S=$(date +%s); timeout 600 $RUN bench.py --impl reference --gpus $N --steps 5 --warmup 3 > $O/m2_ref_$N.log 2> $O/m2_ref_$N.err; echo "reference arm rc=$? wall=$(( $(date +%s) - S ))s"
S=$(date +%s); timeout 600 $RUN bench.py --gpus $N --steps 5 --warmup 3 > $O/m2_bench_$N.log 2> $O/m2_bench_$N.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))s"
tail -c 600 $O/m2_bench_$N.err
python - $O/m2_bench_$N.log $O/m2_ref_$N.log <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("ours N=%d: value %.4g ms_per_step %.3f e2e %.4g (%.3f ms) scaling %s" % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['scaling']))
for k in ('c4_strong', 'c3_strong', 'ot_sharded'):
    if k in d: print(k, json.dumps(d[k])[:600])
r=[l for l in open(sys.argv[2]) if l.startswith('{')]
print("ref:", r[-1][:400] if r else "no line")
PY
