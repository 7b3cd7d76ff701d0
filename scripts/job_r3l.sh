#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_multibox_gpu.py tests/test_sphr_velocity_gpu.py tests/test_eos.py -q -k "velpred or premac or firstdt or velocity" > gpurun_out/r3l_pytest.log 2>&1; tail -6 gpurun_out/r3l_pytest.log
PERF_ONLY=periodic,1 python scripts/perf_episodes.py 256 > gpurun_out/r3l_perf256.log 2>&1; grep "n=256" gpurun_out/r3l_perf256.log
