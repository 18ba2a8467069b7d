#!/bin/bash
# N-GPU bench line (the driver's SCALE run at this N)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-8}
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
URE_BENCH_DEBUG=1 timeout 1200 $RUN bench.py --gpus $N --steps 5 --warmup 3 > $O/r2n${N}_bench.log 2> $O/r2n${N}_bench.err
echo "rc=$?"
tail -c 1200 $O/r2n${N}_bench.err; tail -c 200 $O/r2n${N}_bench.log
timeout 600 $RUN bench.py --gpus $N --impl reference --steps 1 --warmup 0 > $O/r2n${N}_ref.log 2> $O/r2n${N}_ref.err
tail -c 300 $O/r2n${N}_ref.log
