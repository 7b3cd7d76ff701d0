#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/perf_edge2.py 64 > gpurun_out/r2c_perf64.log 2>&1; tail -4 gpurun_out/r2c_perf64.log
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "fused or edge" > gpurun_out/r2c_pytest.log 2>&1
tail -12 gpurun_out/r2c_pytest.log
timeout 300 python scripts/perf_edge2.py 256 > gpurun_out/r2c_perf_edge2.log 2>&1
tail -8 gpurun_out/r2c_perf_edge2.log
