#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_parity_gpu.py -x -q -k "test_fused_edge_2d and march and moving" > gpurun_out/r3s_race_march.log 2>&1
echo "== racecheck march: $(grep -E 'passed|failed' gpurun_out/r3s_race_march.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r3s_race_march.log | tail -1)"
