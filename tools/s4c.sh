#!/bin/bash
# GPU session 4c: halo convs for 64/128-wide tiles (decoder), vectorised diag kernel, smem direct conv; ProPainter and LAMA/DBNet regression.
set -u
mkdir -p gpurun_out
sum=gpurun_out/s4c_summary.txt; : > $sum
t() { local secs=$1 name=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/s4c_$name.log 2> gpurun_out/s4c_$name.err; local rc=$?
      echo "=== $name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/s4c_$name.log | cut -c1-250)" | tee -a $sum; [ $rc -ne 0 ] && tail -n 6 gpurun_out/s4c_$name.err | cut -c1-300 | tee -a $sum; return $rc; }
t 300 ops python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120
t 300 bench python bench.py --steps 6 --warmup 3 --no-cpu
t 900 sttn python -m pytest tests/test_gpu_sttn.py tests/test_gpu_sttn_det.py -m gpu -q --timeout 600
t 600 pp python -m pytest tests/test_gpu_raft.py tests/test_gpu_zz_pp_ops.py tests/test_gpu_dbnet.py -m gpu -q --timeout 400
t 300 bench_pp python bench.py --workload propainter --steps 2 --warmup 1 --pp-frames 40 --no-cpu
VSR_RT_CONV_HALO=1 t 600 rt_halo python -m pytest tests/test_gpu_lama.py tests/test_gpu_dbnet.py tests/test_gpu_raft.py -m gpu -q --timeout 400
VSR_RT_CONV_HALO=1 t 300 bench_pp_halo python bench.py --workload propainter --steps 2 --warmup 1 --pp-frames 40 --no-cpu
VSR_RT_CONV_HALO=1 t 300 lama_halo python bench.py --workload lama --steps 4 --warmup 3 --no-cpu
t 300 lama python bench.py --workload lama --steps 4 --warmup 3 --no-cpu
timeout 500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm2_kernel -s 40 -c 6 -o gpurun_out/s4c_ncu_gemm2 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/s4c_ncu.log 2>&1
echo "ncu rc=$?" | tee -a $sum
cat $sum
