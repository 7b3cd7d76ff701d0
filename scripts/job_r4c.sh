#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_multibox_gpu.py tests/test_sphr_velocity_gpu.py -q -m gpu -k "velocity or minmax or premac" > gpurun_out/r4c_pytest.log 2>&1; tail -4 gpurun_out/r4c_pytest.log
for pp in 1 2; do
PERF_ONLY=periodic,$pp timeout 600 python scripts/perf_episodes.py 256 > gpurun_out/r4c_perf256_ppm$pp.log 2>&1; grep -i "advance" gpurun_out/r4c_perf256_ppm$pp.log
done
PERF_ONLY=walls,2 timeout 600 python scripts/perf_episodes.py 256 > gpurun_out/r4c_perf256_walls2.log 2>&1; grep -i "advance" gpurun_out/r4c_perf256_walls2.log
