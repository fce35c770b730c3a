#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <script>   -- retries while the pod answers busy (exit 3), nothing is charged for those
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2"
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  sleep 90
done
exit 3
