#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_multibox_gpu.py -q -m gpu -k "premac or mkutrans or minmax or estdt" > gpurun_out/r4b_pytest.log 2>&1; tail -4 gpurun_out/r4b_pytest.log
for pp in 1 2; do
PERF_ONLY=periodic,$pp timeout 600 python scripts/perf_episodes.py 256 > gpurun_out/r4b_perf256_ppm$pp.log 2>&1; grep -i "premac\|velocity" gpurun_out/r4b_perf256_ppm$pp.log
done
