#!/bin/bash
mkdir -p gpurun_out
python scripts/perf_edge2.py 256 > gpurun_out/r2b_perf_edge2.log 2>&1
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "fused or edge" > gpurun_out/r2b_pytest.log 2>&1
tail -8 gpurun_out/r2b_perf_edge2.log; tail -3 gpurun_out/r2b_pytest.log
