#!/bin/bash
mkdir -p gpurun_out
R="timeout 300 compute-sanitizer --tool racecheck --print-limit 40 python -m pytest -x -q"
$R tests/test_parity_gpu.py -k "test_fused_edge_ragged_boxes and fast and shape3" > gpurun_out/r4t_race_edge2.log 2>&1
echo "race fallback kernel, odd pitch (33x47x40): $(grep -E 'passed|failed' gpurun_out/r4t_race_edge2.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r4t_race_edge2.log | tail -1)"
$R tests/test_parity_gpu.py -k "test_fused_edge_slow_faces and upwind-first and shape3" > gpurun_out/r4t_race_slow.log 2>&1
echo "race slow faces (45x33x70): $(grep -E 'passed|failed' gpurun_out/r4t_race_slow.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r4t_race_slow.log | tail -1)"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "fused_edge or make_edge_scal" > gpurun_out/r4t_pytest.log 2>&1; tail -2 gpurun_out/r4t_pytest.log
