#!/bin/bash
# usage: tools/gpurun_retry.sh "<gpurun args>" -- retries while the pod answers busy (exit 3 / transient)
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|exit code 3\|no box or slot"; then sleep 150; continue; fi
  break
done
