#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (nothing is charged for those).  usage: gpurun_retry.sh <timeout> [--gpus N] <command>
t=$1; shift
extra=""
if [ "$1" = "--gpus" ]; then extra="--gpus $2"; shift; shift; fi
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t $extra -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then
    sleep 150
    continue
  fi
  echo "$out"
  exit 0
done
echo "$out"
exit 3
