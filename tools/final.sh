#!/bin/bash
# Final verification at HEAD on one B200: build check, smoke, the whole GPU suite, the driver's bench command, a short reference-arm run.
set -u
mkdir -p gpurun_out
sum=gpurun_out/final_summary.txt; : > $sum
t() { local secs=$1 name=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/final_$name.log 2> gpurun_out/final_$name.err; local rc=$?
      echo "=== $name rc=$rc $(( $(date +%s) - t0 ))s :: $(grep -v '^\s*$' gpurun_out/final_$name.log | tail -n 1 | cut -c1-300)" | tee -a $sum; [ $rc -ne 0 ] && tail -n 8 gpurun_out/final_$name.err | cut -c1-300 | tee -a $sum; return $rc; }
git_head=$(cat .git/HEAD 2>/dev/null || echo "snapshot")
echo "head: $git_head" | tee -a $sum
t 200 smoke python -c "import __graft_entry__ as g; g.build(); g.smoke()"
PYTHONFAULTHANDLER=1 t 1100 pytest_gpu python -m pytest tests -m gpu -q --timeout 300
t 400 bench20 python bench.py --gpus 1 --steps 20 --warmup 5
t 200 ref python bench.py --impl reference --gpus 1 --steps 2 --warmup 1
cat $sum
