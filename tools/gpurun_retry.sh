#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (nothing is charged for those): tools/gpurun_retry.sh <timeout> '<command>'
t=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3"; then sleep 90; continue; fi
  break
done
