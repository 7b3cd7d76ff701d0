#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_average.py tests/test_eos.py -q > gpurun_out/r2w_pytest.log 2>&1; tail -30 gpurun_out/r2w_pytest.log
