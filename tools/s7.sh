#!/bin/bash
# single-clip strong scaling at N = 4 (last GPU seconds of round 2)
set -u
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --workload sttn-auto-strong --steps 2 --warmup 1 > gpurun_out/s7_strong4.log 2> gpurun_out/s7_strong4.err
echo "rc=$?"; grep -v '^\s*$' gpurun_out/s7_strong4.log | tail -n 1 | cut -c1-300
