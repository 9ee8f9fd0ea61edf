#!/bin/bash
# First GPU session of the ProPainter path: run on the B200 box as
#   gpurun --timeout 1500 -- 'bash tools/bringup_propainter.sh'
# after removing weights/propainter from .gpurunignore.  Every stage runs under its own timeout and writes its log to gpurun_out/, so that one
# hung or faulting stage neither costs the whole call nor hides the stages before it.  Stages are ordered from the smallest unit to the whole
# pipeline; the bench line comes last and only if the parity tests passed.
set -u
mkdir -p gpurun_out
export VSR_RUN_UNVALIDATED=1
status=0
run() {   # name, seconds, command...
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/bringup_summary.txt
  timeout "$secs" "$@" > "gpurun_out/bringup_$name.log" 2>&1
  local rc=$?
  echo "    rc=$rc  $(tail -n 1 "gpurun_out/bringup_$name.log")" | tee -a gpurun_out/bringup_summary.txt
  [ $rc -ne 0 ] && status=1
  return $rc
}
run validated 600 python -m pytest tests -m gpu -x -q --ignore tests/test_gpu_raft.py --ignore tests/test_gpu_zz_pp_ops.py   # nothing regressed first
run pp_ops 600 python -m pytest tests/test_gpu_zz_pp_ops.py -m gpu -q                                # every ProPainter kernel alone, no weights needed
for t in $(python -m pytest tests/test_gpu_raft.py -m gpu --collect-only -q 2>/dev/null | grep "::"); do
  run "$(echo "$t" | sed 's/.*:://; s/[^A-Za-z0-9_]/_/g')" 300 python -m pytest "$t" -x -q -s
done
if [ $status -eq 0 ]; then
  run bench_propainter 600 python bench.py --workload propainter --steps 3 --warmup 3
  run launches_propainter 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_propainter.csv \
      python bench.py --workload propainter --steps 1 --warmup 3 --pp-frames 12 --no-cpu
fi
cat gpurun_out/bringup_summary.txt
exit $status
