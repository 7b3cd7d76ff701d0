#!/bin/bash
mkdir -p gpurun_out
PERF_ONLY=periodic,1 python scripts/perf_episodes.py 256 > gpurun_out/r2v_perf256.log 2>&1; cat gpurun_out/r2v_perf256.log | tail -12
PERF_ONLY=periodic,1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2v_launches_premac.csv python scripts/perf_episodes.py 256 > gpurun_out/r2v_ncu.log 2>&1; echo "ncu rc=$?"
