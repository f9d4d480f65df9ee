#!/bin/bash
# usage: gpuretry.sh <timeout> <command...>   - retries while the pod answers busy/transient
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -4
  if echo "$out" | grep -q "status=transient\|exit code 3\|busy"; then sleep 90; continue; fi
  break
done
