#!/bin/bash
# usage: scripts/gpu_retry.sh <logfile> [gpurun args...] -- '<command>'
# Re-submits a gpurun call while the pod answers "transient"/busy (nothing is charged for those).
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$log" || [ $rc -eq 3 ]; then sleep 120; continue; fi
  exit $rc
done
exit 3
