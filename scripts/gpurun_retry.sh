#!/bin/bash
# usage: bash scripts/gpurun_retry.sh <timeout> <command...>   -- retries while the pod answers busy / transient (nothing charged)
t=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  echo "$out"
  if echo "$out" | grep -q "status=transient\|status=busy\|retry in a few minutes"; then sleep 120; continue; fi
  break
done
