#!/bin/bash
# usage: tools/gpu_retry.sh TIMEOUT 'command'   — retries while the pod answers busy/transient (nothing is charged for those)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@"
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' /root/repo/gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  echo "[gpu_retry] busy (attempt $i), sleeping 90 s"; sleep 90
done
exit 3
