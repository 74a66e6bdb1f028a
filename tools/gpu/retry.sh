#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit 3): tools/gpu/retry.sh <timeout-seconds> <script>
for i in $(seq 1 40); do
  gpurun --timeout "$1" -- "bash $2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
