#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multibox_gpu.py -x -q > gpurun_out/r2q_pytest.log 2>&1; tail -15 gpurun_out/r2q_pytest.log
