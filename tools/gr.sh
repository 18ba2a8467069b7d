#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3): tools/gr.sh <timeout_s> <out_file> <command...>
T=$1; OUT=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $OUT 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $OUT; then echo "rc=$rc" >> $OUT; exit $rc; fi
  sleep 90
done
echo "rc=3 (gave up)" >> $OUT; exit 3
