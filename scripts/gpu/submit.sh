#!/bin/bash
# submit.sh <name> <timeout> [gpurun args...] : run scripts/gpu/<name>.sh on the GPU box, retrying while the pod is busy
name=$1; limit=$2; shift 2
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$limit" "$@" -- "bash scripts/gpu/$name.sh" > gpurun_out/call_$name.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/call_$name.log; then break; fi
  echo "attempt $attempt: busy, retrying in 150 s" >> gpurun_out/call_$name.retries
  sleep 150
done
tail -60 gpurun_out/call_$name.log | cut -c1-700
