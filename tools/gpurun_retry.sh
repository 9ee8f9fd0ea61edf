#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3, nothing charged):
#   gpurun_retry.sh <timeout> <logfile> [--gpus N] <command>
t=$1; log=$2; shift 2
opts=()
if [ "${1:-}" = "--gpus" ]; then opts=(--gpus "$2"); shift 2; fi
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "${opts[@]}" --timeout "$t" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 150
done
exit 3
