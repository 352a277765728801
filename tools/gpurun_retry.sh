#!/bin/bash
# tools/gpurun_retry.sh TIMEOUT 'command'  — re-queues while the pod answers "busy / transient / refused" (nothing charged)
T=$1; shift
for i in $(seq 1 200); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|no box\|another call"; then sleep 8; continue; fi
  if [ $rc -eq 2 ] || [ $rc -eq 3 ]; then sleep 8; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up"
