#!/bin/bash
# GPU session 3: coalesced conv epilogue + 8 epilogue warps; whole validated suite (the conv epilogue is shared by every network);
# config-size goldens; cluster-of-4 weight multicast A/B; ProPainter stage breakdown; the unmodified reference on the box's CPU.
set -u
mkdir -p gpurun_out
sum=gpurun_out/s3_summary.txt; : > $sum
t() { local secs=$1 name=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/s3_$name.log 2> gpurun_out/s3_$name.err; local rc=$?
      echo "=== $name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/s3_$name.log | cut -c1-250)" | tee -a $sum; [ $rc -ne 0 ] && tail -n 12 gpurun_out/s3_$name.err | cut -c1-300 | tee -a $sum; return $rc; }
t 300 conv python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120
t 300 bench python bench.py --steps 6 --warmup 3 --no-cpu
VSR_CONV_CLUSTER=4 VSR_DEBUG_CLUSTERS=1 t 200 conv_cl4 python -m pytest tests/test_gpu_ops.py -m gpu -q -k conv_igemm --timeout 60
if [ $? -eq 0 ]; then VSR_CONV_CLUSTER=4 t 300 bench_cl4 python bench.py --steps 6 --warmup 3 --no-cpu; fi
t 900 suite python -m pytest tests -m gpu -q --timeout 600 --ignore tests/test_gpu_ops.py --ignore tests/test_gpu_zz_pp_ops.py -x
t 300 bench_det python bench.py --workload sttn-det --steps 4 --warmup 3
t 300 bench_pp python bench.py --workload propainter --steps 2 --warmup 1 --pp-frames 40 --no-cpu
t 300 ref python bench.py --impl reference --steps 2 --warmup 1
t 300 lama512 python bench.py --workload lama512 --steps 3 --warmup 3
cat $sum
