#!/bin/bash
# round-2 multi-GPU pass (N = $1 GPUs of one box)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
O=gpurun_out
nvidia-smi topo -m > $O/r2mm_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $O/r2mm_multi_tests.log 2>&1
tail -15 $O/r2mm_multi_tests.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
URE_BENCH_DEBUG=1 timeout 900 $RUN bench.py --gpus $N --steps 5 --warmup 3 > $O/r2mm_bench_$N.log 2> $O/r2mm_bench_$N.err
tail -c 1500 $O/r2mm_bench_$N.err; tail -c 3000 $O/r2mm_bench_$N.log




