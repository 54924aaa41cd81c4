#!/bin/bash
# usage: scripts/gpurun_retry.sh <gpus> <timeout> <command-file>   - retries while the pod is busy (exit code 3)
gpus=$1; to=$2; cmdfile=$3
for i in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then
    /usr/local/graft/bin/gpurun --timeout $to -- "$(cat $cmdfile)"
  else
    /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$(cat $cmdfile)"
  fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 100
done
exit 3
