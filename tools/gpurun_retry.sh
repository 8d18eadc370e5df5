#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>   — retries while the pod answers "transient"/busy
T=$1; shift
for attempt in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -80
  if echo "$out" | grep -q "status=transient\|exit code 3\|rc=3"; then sleep 120; continue; fi
  break
done
