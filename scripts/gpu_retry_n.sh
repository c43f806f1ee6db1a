#!/bin/bash
# usage: scripts/gpu_retry_n.sh <gpus> <timeout_s> <logfile> '<command>'   -- like gpu_retry.sh on N GPUs of one box
N=$1; T=$2; LOG=$3; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $N --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then echo "done rc=$rc" >> $LOG; exit $rc; fi
  sleep 45
done
echo "gave up" >> $LOG
