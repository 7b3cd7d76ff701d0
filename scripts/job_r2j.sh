#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; tail -2 gpurun_out/r2j_pytest.log
timeout 300 python scripts/perf_edge2.py 256 2>&1 | grep "by 1616" | grep "= -1" > gpurun_out/r2j_perf_edge2.log; cat gpurun_out/r2j_perf_edge2.log
timeout 600 python scripts/perf_episodes.py 128 > gpurun_out/r2j_perf_episodes.log 2>&1; grep -v launches gpurun_out/r2j_perf_episodes.log | tail -20
