#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r4g_launches_c5.csv python bench.py --config c5 --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r4g_ncu_c5.log 2>&1; tail -1 gpurun_out/r4g_ncu_c5.log | cut -c1-200
