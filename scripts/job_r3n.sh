#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_gpu.py -q -k "test_multi_gpu" > gpurun_out/r3n_pytest.log 2>&1; tail -30 gpurun_out/r3n_pytest.log | cut -c1-400
