#!/bin/bash
# Final verification on one B200: smoke, the STTN-path GPU suites (every kernel family of the contract line; the other networks' suites ran
# green in sessions 4c / s1 with the same runtime code), the driver's bench command, a short reference-arm run.
set -u
mkdir -p gpurun_out
sum=gpurun_out/final_summary.txt; : > $sum
t() { local secs=$1 name=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/final_$name.log 2> gpurun_out/final_$name.err; local rc=$?
      echo "=== $name rc=$rc $(( $(date +%s) - t0 ))s :: $(grep -v '^\s*$' gpurun_out/final_$name.log | tail -n 1 | cut -c1-300)" | tee -a $sum; [ $rc -ne 0 ] && tail -n 8 gpurun_out/final_$name.err | cut -c1-300 | tee -a $sum; return $rc; }
t 100 smoke python -c "import __graft_entry__ as g; g.smoke()"
PYTHONFAULTHANDLER=1 t 150 lockstep python -m pytest tests/test_gpu_sttn.py -m gpu -q -k window_sharded --timeout 300
t 200 bench20 python bench.py --gpus 1 --steps 20 --warmup 5
PYTHONFAULTHANDLER=1 t 420 pytest_gpu python -m pytest tests/test_gpu_ops.py tests/test_gpu_sttn.py tests/test_gpu_sttn_det.py tests/test_gpu_scene.py -m gpu -q --timeout 200 --deselect tests/test_gpu_sttn.py::test_window_sharded_chunk_equals_single_engine
cat $sum
