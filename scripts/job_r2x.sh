#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_sphr_velocity_gpu.py -q -k "mkutrans or velpred or premac or velocity_advance or vel" > gpurun_out/r2x_pytest.log 2>&1; tail -8 gpurun_out/r2x_pytest.log
PERF_ONLY=periodic,1 python scripts/perf_episodes.py 256 > gpurun_out/r2x_perf256.log 2>&1; grep "n=256" gpurun_out/r2x_perf256.log
