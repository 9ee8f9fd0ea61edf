#!/bin/bash
# GPU session 5 (2 GPUs): NCCL paths — window-sharded chunk bit-identity, single-clip strong scaling, sharded config 4, weak-scaling contract line.
set -u
mkdir -p gpurun_out
sum=gpurun_out/s5_summary.txt; : > $sum
N=${1:-2}
t() { local secs=$1 name=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/s5_$name.log 2> gpurun_out/s5_$name.err; local rc=$?
      echo "=== $name rc=$rc $(( $(date +%s) - t0 ))s :: $(grep -v '^\s*$' gpurun_out/s5_$name.log | tail -n 1 | cut -c1-300)" | tee -a $sum; [ $rc -ne 0 ] && tail -n 8 gpurun_out/s5_$name.err | cut -c1-300 | tee -a $sum; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PYTHONFAULTHANDLER=1 t 300 lockstep python -m pytest tests/test_gpu_sttn.py -m gpu -q -k "window_sharded" --timeout 120
t 400 sharded_check $TR --master-port 29511 tools/run_sharded_check.py
t 300 strong $TR --master-port 29512 bench.py --gpus $N --workload sttn-auto-strong --steps 2 --warmup 1
t 300 weak $TR --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3
t 400 config4 $TR --master-port 29514 bench.py --gpus $N --workload config4 --steps 2 --warmup 1
nvidia-smi topo -m > gpurun_out/s5_topo.txt 2>&1
cat $sum
