#!/bin/bash
# retries a gpurun call while the pod answers "busy / transient" (nothing is charged for those)
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient\|nothing was charged" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
cat /tmp/gpurun_last.log | tail -${TAILN:-40}
exit $rc
