#!/bin/bash
# scratch job run on the GPU box by gpurun (edited per experiment)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 120 python scripts/perf_2d.py 4096 > gpurun_out/q_perf_2d.log 2>&1
tail -3 gpurun_out/q_perf_2d.log
