#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> [--gpus N] -- <command>   (retries while the pod answers busy / transient)
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|rc=3\|no box\|answers busy" $LOG && ! grep -q "exit code" $LOG; then
    if [ $rc -eq 3 ] || grep -q "status=transient" $LOG; then sleep 150; continue; fi
  fi
  break
done
echo "gpurun_retry done rc=$rc attempt=$i" >> $LOG
