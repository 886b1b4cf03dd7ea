#!/bin/bash
# usage: scripts/gpurun_retry.sh <out.txt> <timeout> <command...>   - retries while the pod answers "busy/transient"
out=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $out 2>&1
  if ! grep -q "status=transient\|nothing was charged" $out; then break; fi
  sleep 150
done
