#!/bin/bash
# Last GPU seconds of round 2 (2 GPUs): the window-sharded chunk under NCCL after the work-buffer fix, then single-clip strong scaling at N = 2.
set -u
mkdir -p gpurun_out
sum=gpurun_out/s6_summary.txt; : > $sum
t() { local secs=$1 name=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/s6_$name.log 2> gpurun_out/s6_$name.err; local rc=$?
      echo "=== $name rc=$rc $(( $(date +%s) - t0 ))s :: $(grep -v '^\s*$' gpurun_out/s6_$name.log | tail -n 1 | cut -c1-300)" | tee -a $sum; [ $rc -ne 0 ] && tail -n 12 gpurun_out/s6_$name.err | cut -c1-300 | tee -a $sum; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
PYTHONFAULTHANDLER=1 t 110 sharded_check $TR --master-port 29511 tools/run_sharded_check.py
PYTHONFAULTHANDLER=1 t 90 strong $TR --master-port 29512 bench.py --gpus 2 --workload sttn-auto-strong --steps 2 --warmup 1
cat $sum
