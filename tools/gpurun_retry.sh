#!/bin/bash
# retry a gpurun call while the pod answers "transient / busy" (nothing is charged for those)
# usage: tools/gpurun_retry.sh <timeout> <command...>
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" > /tmp/gpurun_last.log 2>&1
  if grep -q "status=transient\|exit code 3\|rc=3" /tmp/gpurun_last.log && ! grep -q "charged=[1-9]" /tmp/gpurun_last.log; then
    echo "[retry $i] transient, sleeping" >> /tmp/gpurun_retry.log; sleep 100; continue
  fi
  break
done
tail -80 /tmp/gpurun_last.log
