#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/perf_2d.py 4096 > gpurun_out/q_perf_2d.log 2>&1
tail -5 gpurun_out/q_perf_2d.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/q_launches_2d.csv python scripts/perf_2d.py 4096 > /dev/null 2>&1
python scripts/ncu_summary.py launches gpurun_out/q_launches_2d.csv 39
