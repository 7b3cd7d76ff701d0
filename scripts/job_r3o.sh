#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --selftest > gpurun_out/r3o_selftest.log 2>&1; tail -1 gpurun_out/r3o_selftest.log | cut -c1-1500
grep -n "firstdt\|MISMATCH\|Error" gpurun_out/r3o_selftest.log | head
