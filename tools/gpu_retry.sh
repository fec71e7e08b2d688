#!/bin/bash
# usage: tools/gpu_retry.sh TIMEOUT 'command'   -- retries a gpurun call while the pod answers busy / transient
t=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  echo "$out" | tail -120
  if echo "$out" | grep -q "status=transient\|rc=3\|status=busy"; then sleep 150; continue; fi
  break
done
