#!/bin/bash
# 2-GPU job: scaling bench lines, multi-rank selftest, strong-scaling configs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29611 bench.py --gpus 2 > gpurun_out/r2m_bench_c2_n2.log 2>&1; tail -1 gpurun_out/r2m_bench_c2_n2.log | cut -c1-400
$TR --master-port 29612 bench.py --gpus 2 --selftest > gpurun_out/r2m_selftest_n2.log 2>&1; tail -1 gpurun_out/r2m_selftest_n2.log | cut -c1-1200
$TR --master-port 29613 bench.py --gpus 2 --config c4 --steps 5 > gpurun_out/r2m_bench_c4_n2.log 2>&1; tail -1 gpurun_out/r2m_bench_c4_n2.log | cut -c1-400
$TR --master-port 29614 bench.py --gpus 2 --config c5 --steps 3 > gpurun_out/r2m_bench_c5_n2.log 2>&1; tail -1 gpurun_out/r2m_bench_c5_n2.log | cut -c1-400
