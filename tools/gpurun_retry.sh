#!/bin/bash
# tools/gpurun_retry.sh [gpurun args...] -- retries a gpurun call while the pod answers "busy / transient" (nothing charged)
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|nothing was charged — retry\|no box or slot"; then sleep 120; continue; fi
  exit $rc
done
exit 3
