#!/bin/bash
# GPU session 2: bring-up of the haloed-tile conv kernel (conv_halo.cuh).  Decides the descriptor base-offset question by test, then measures.
set -u
mkdir -p gpurun_out
sum=gpurun_out/s2_summary.txt; : > $sum
t() { local name=$1; shift; timeout 300 "$@" > gpurun_out/s2_$name.log 2>&1; local rc=$?; echo "=== $name rc=$rc :: $(tail -n 1 gpurun_out/s2_$name.log | cut -c1-250)" | tee -a $sum; return $rc; }
mode=off
if VSR_CONV_HALO_BASEOFF=1 t conv_baseoff1 python -m pytest tests/test_gpu_ops.py -m gpu -q -k conv_igemm --timeout 120; then mode=base1
elif VSR_CONV_HALO_BASEOFF=0 t conv_baseoff0 python -m pytest tests/test_gpu_ops.py -m gpu -q -k conv_igemm --timeout 120; then mode=base0; fi
echo "halo mode: $mode" | tee -a $sum
case $mode in base1) ;; base0) export VSR_CONV_HALO_BASEOFF=0 ;; off) export VSR_CONV_HALO=0 ;; esac
t sttn python -m pytest tests/test_gpu_sttn.py tests/test_gpu_sttn_det.py -m gpu -q --timeout 300
t bench_halo python bench.py --steps 6 --warmup 3 --no-cpu
VSR_CONV_HALO=0 t bench_nohalo python bench.py --steps 6 --warmup 3 --no-cpu
t bench_det python bench.py --workload sttn-det --steps 4 --warmup 3
if [ $mode != off ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 40 -c 2 -o gpurun_out/s2_ncu_conv_halo -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/s2_ncu_conv.log 2>&1
  echo "ncu conv rc=$?" | tee -a $sum
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/s2_launches.log 2>&1
echo "launch list rc=$?" | tee -a $sum
cat $sum
