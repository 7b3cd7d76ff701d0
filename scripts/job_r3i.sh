#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multibox_gpu.py -q > gpurun_out/r3i_pytest.log 2>&1; tail -25 gpurun_out/r3i_pytest.log
