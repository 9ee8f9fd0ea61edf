#!/bin/bash
# GPU session 4b: direct attention with the self-logit row shift; suite; benches; ncu of the QKV / score / PV kernels; ProPainter launch list.
set -u
mkdir -p gpurun_out
sum=gpurun_out/s4b_summary.txt; : > $sum
t() { local secs=$1 name=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/s4b_$name.log 2> gpurun_out/s4b_$name.err; local rc=$?
      echo "=== $name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/s4b_$name.log | cut -c1-250)" | tee -a $sum; [ $rc -ne 0 ] && tail -n 6 gpurun_out/s4b_$name.err | cut -c1-300 | tee -a $sum; return $rc; }
t 300 ops python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120
t 300 bench python bench.py --steps 6 --warmup 3 --no-cpu
t 900 sttn python -m pytest tests/test_gpu_sttn.py tests/test_gpu_sttn_det.py tests/test_gpu_scene.py -m gpu -q --timeout 600
t 300 bench_det python bench.py --workload sttn-det --steps 4 --warmup 3
for k in Conv2Policy Score2Policy PV2Policy; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 2 -o gpurun_out/s4b_ncu_$k -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/s4b_ncu_$k.log 2>&1
  echo "ncu $k rc=$?" | tee -a $sum
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/s4b_pp_launches.csv python bench.py --workload propainter --steps 1 --warmup 1 --pp-frames 20 --no-cpu > gpurun_out/s4b_pp_launches.log 2>&1
echo "pp launch list rc=$?" | tee -a $sum
cat $sum
