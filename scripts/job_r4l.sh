#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_eos.py tests/test_sphr_velocity_gpu.py tests/test_average.py -q -m gpu > gpurun_out/r4l_pytest.log 2>&1; tail -6 gpurun_out/r4l_pytest.log
