#!/bin/bash
# One GPU session = one gpurun call: every stage under its own timeout, its log under gpurun_out/<tag>_<stage>.log, a summary at the end.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh <tag> <stage>...'
# Stages: validated | sttn | bench | bench20 | ppops | pp | pp_bench | det | lama | dbnet | ref
set -u
tag=$1; shift
mkdir -p gpurun_out
sum=gpurun_out/${tag}_summary.txt
: > "$sum"
run() {   # name, seconds, command...
  local name=$1 secs=$2; shift 2
  local t0=$(date +%s)
  timeout "$secs" "$@" > "gpurun_out/${tag}_${name}.log" 2> "gpurun_out/${tag}_${name}.err"
  local rc=$?
  echo "=== $name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 "gpurun_out/${tag}_${name}.log" | cut -c1-300)" | tee -a "$sum"
  [ $rc -ne 0 ] && tail -n 15 "gpurun_out/${tag}_${name}.err" | cut -c1-400 | tee -a "$sum"
  return $rc
}
for stage in "$@"; do
  case $stage in
    validated) run validated 900 python -m pytest tests -m gpu -q --timeout 600 --ignore tests/test_gpu_raft.py --ignore tests/test_gpu_zz_pp_ops.py ;;
    sttn)      run sttn 600 python -m pytest tests/test_gpu_sttn.py tests/test_gpu_ops.py -m gpu -q --timeout 300 ;;
    bench)     run bench 600 python bench.py --steps 6 --warmup 3 ;;
    bench20)   run bench20 900 python bench.py --gpus 1 --steps 20 --warmup 5 ;;
    ref)       run ref 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ;;
    ppops)     run ppops 600 python -m pytest tests/test_gpu_zz_pp_ops.py -m gpu -q --timeout 300 ;;
    pp)        for t in test_raft_flows_vs_oracle test_image_propagation_vs_oracle test_flow_completion_vs_oracle test_propainter_pipeline_vs_reference_frames test_batched_detection_equals_single_frames; do
                 run "pp_$t" 400 python -m pytest "tests/test_gpu_raft.py::$t" -x -q -s --timeout 380
               done ;;
    pp_bench)  run pp_bench 900 python bench.py --workload propainter --steps 2 --warmup 1 --pp-frames 20 ;;
    det)       run det 600 python bench.py --workload sttn-det --steps 6 --warmup 3 ;;
    lama)      run lama 600 python bench.py --workload lama --steps 4 --warmup 3 ;;
    dbnet)     run dbnet 600 python bench.py --workload dbnet --steps 4 --warmup 3 ;;
    *)         run "$stage" 900 bash -c "$stage" ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader >> "$sum" 2>&1
cat "$sum"
exit 0
