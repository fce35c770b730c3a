#!/bin/bash
# usage: tools/gpurun_retry2.sh <timeout_s> <script>   -- 2-GPU variant of gpurun_retry.sh
for i in $(seq 1 10); do
  /usr/local/graft/bin/gpurun --gpus 2 --timeout "$1" -- "bash $2"
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  sleep 90
done
exit 3
