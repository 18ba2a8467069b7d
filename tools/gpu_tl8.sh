#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-8}
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
URE_BENCH_TIMELINE=1 URE_BENCH_DEBUG=1 timeout 900 $RUN bench.py --gpus $N --steps 5 --warmup 3 --no-extra --no-cpu > $O/r2tl${N}_bench.log 2> $O/r2tl${N}_bench.err
grep "\[timeline\]" $O/r2tl${N}_bench.err | tail -45 | cut -c1-150
tail -c 400 $O/r2tl${N}_bench.log
