#!/bin/bash
# scratch job run on the GPU box by gpurun (edited per experiment)
mkdir -p gpurun_out
export BENCH_WATCHDOG=70
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 2 --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/m_dbg_n2.log 2>&1
grep -v "^W1017\|^\*\*\*\|OMP_NUM" gpurun_out/m_dbg_n2.log | head -60 | cut -c1-400
