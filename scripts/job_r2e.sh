#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "fused or edge" > gpurun_out/r2e_pytest.log 2>&1
tail -2 gpurun_out/r2e_pytest.log
timeout 300 python scripts/perf_edge2.py 256 > gpurun_out/r2e_perf_edge2.log 2>&1
grep "by 1616" gpurun_out/r2e_perf_edge2.log
ncu --set full --clock-control none --import-source on -k regex:k_fused_edge3 -s 1 -c 1 -o gpurun_out/prof_fused3_e python scripts/one_edge.py 256 1 3 > gpurun_out/r2e_ncu.log 2>&1; tail -1 gpurun_out/r2e_ncu.log
