#!/bin/bash
# GPU session 4d: TMA-store epilogue (UTMASTG) for fp16-only conv outputs; the sttn suite test by test (a test hung in session 4c).
set -u
mkdir -p gpurun_out
sum=gpurun_out/s4d_summary.txt; : > $sum
t() { local secs=$1 name=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/s4d_$name.log 2> gpurun_out/s4d_$name.err; local rc=$?
      echo "=== $name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/s4d_$name.log | cut -c1-250)" | tee -a $sum; [ $rc -ne 0 ] && tail -n 6 gpurun_out/s4d_$name.err | cut -c1-300 | tee -a $sum; return $rc; }
t 200 ops python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 100
if [ $? -ne 0 ]; then export VSR_CONV_TMA_STORE=0; echo "TMA store OFF for the rest" | tee -a $sum; VSR_CONV_TMA_STORE=0 t 200 ops_notma python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 100; fi
t 200 bench python bench.py --steps 6 --warmup 3 --no-cpu
VSR_CONV_TMA_STORE=0 t 200 bench_notma python bench.py --steps 6 --warmup 3 --no-cpu
PYTHONFAULTHANDLER=1 t 700 sttn python -m pytest tests/test_gpu_sttn.py tests/test_gpu_sttn_det.py tests/test_gpu_scene.py -m gpu -v --timeout 150 -x
t 200 bench_det python bench.py --workload sttn-det --steps 4 --warmup 3
cat $sum
