#!/bin/bash
# gpurun with retries while the pod has no free slot: tools/gr.sh <timeout_s> <out_file> <command...>   (GPUS=N for --gpus N)
T=$1; OUT=$2; shift 2
G=""; if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@" > $OUT 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $OUT && ! grep -q "status=busy" $OUT; then echo "rc=$rc" >> $OUT; exit $rc; fi
  sleep 60
done
echo "rc=3 (gave up)" >> $OUT; exit 3
