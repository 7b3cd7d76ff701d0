#!/bin/bash
mkdir -p gpurun_out
export BENCH_WATCHDOG=200
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/t_bench_n2.log 2>&1
grep '^{' gpurun_out/t_bench_n2.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=2', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'])"
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 2>&1 | grep '^{' | cut -c1-250
