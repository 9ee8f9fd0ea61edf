#!/bin/bash
# GPU session 4: direct attention (fp16), coalesced epilogue, window sharding (1-GPU engines), scene cuts; whole suite; benches.
set -u
mkdir -p gpurun_out
sum=gpurun_out/s4_summary.txt; : > $sum
t() { local secs=$1 name=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/s4_$name.log 2> gpurun_out/s4_$name.err; local rc=$?
      echo "=== $name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/s4_$name.log | cut -c1-250)" | tee -a $sum; [ $rc -ne 0 ] && tail -n 6 gpurun_out/s4_$name.err | cut -c1-300 | tee -a $sum; return $rc; }
t 300 ops python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120
if [ $? -ne 0 ]; then
  VSR_ATTN_DIRECT=0 t 300 ops_nodirect python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120 && export VSR_ATTN_DIRECT=0 && echo "direct attention OFF for the rest" | tee -a $sum
fi
t 300 bench python bench.py --steps 6 --warmup 3 --no-cpu
VSR_ATTN_DIRECT=0 t 300 bench_nodirect python bench.py --steps 6 --warmup 3 --no-cpu
VSR_CONV_CLUSTER=4 t 300 bench_cl4 python bench.py --steps 6 --warmup 3 --no-cpu
t 1200 suite python -m pytest tests -m gpu -q --timeout 600 --ignore tests/test_gpu_ops.py --ignore tests/test_gpu_zz_pp_ops.py
t 300 bench_det python bench.py --workload sttn-det --steps 4 --warmup 3
t 300 strong1 python bench.py --workload sttn-auto-strong --steps 2 --warmup 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_halo|Score2|PV2|Conv2Policy" -s 60 -c 5 -o gpurun_out/s4_ncu -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/s4_ncu.log 2>&1
echo "ncu rc=$?" | tee -a $sum
cat $sum
