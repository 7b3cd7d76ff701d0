#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python scripts/perf_2d.py 4096 2>&1 | grep "density_advance" | cut -c1-100
