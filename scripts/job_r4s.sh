#!/bin/bash
# compute-sanitizer on the shared-memory tables added in the second session (slopes / edge values shared through
# shared memory in the 3-D and 2-D kernels), the fused premac face kernel and the paired-stream tile split
mkdir -p gpurun_out
R="timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest -x -q"
M="timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest -x -q"
$R tests/test_parity_gpu.py -k "test_fused_edge_ragged_boxes and fast and shape3" > gpurun_out/r4s_race_edge3.log 2>&1
echo "race edge3 (ppm 0/1/2, periodic/walls/inout, 33x47x40): $(grep -E 'passed|failed' gpurun_out/r4s_race_edge3.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r4s_race_edge3.log | tail -1)"
$R tests/test_parity_gpu.py -k "thin_end_chunks and 1-1-shape0" > gpurun_out/r4s_race_thin.log 2>&1
echo "race thin end chunks + pair streams: $(grep -E 'passed|failed' gpurun_out/r4s_race_thin.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r4s_race_thin.log | tail -1)"
$R tests/test_parity_gpu.py -k "test_density_advance and fast and 3-" > gpurun_out/r4s_race_episode.log 2>&1
echo "race density_advance episodes: $(grep -E 'passed|failed' gpurun_out/r4s_race_episode.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r4s_race_episode.log | tail -1)"
$R tests/test_parity_gpu.py -k "test_fused_edge_2d and 32x16 and moving" > gpurun_out/r4s_race_2d.log 2>&1
echo "race 2d: $(grep -E 'passed|failed' gpurun_out/r4s_race_2d.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r4s_race_2d.log | tail -1)"
$R tests/test_parity_gpu.py -k "test_advance_premac_fused_and_staged_agree" > gpurun_out/r4s_race_premac.log 2>&1
echo "race premac fused: $(grep -E 'passed|failed' gpurun_out/r4s_race_premac.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r4s_race_premac.log | tail -1)"
$M tests/test_parity_gpu.py -k "(thin_end_chunks and 1-1-shape0) or test_advance_premac_fused_and_staged_agree or minmax" > gpurun_out/r4s_mem.log 2>&1
echo "memcheck thin / premac / minmax: $(grep -E 'passed|failed' gpurun_out/r4s_mem.log | tail -1) | $(grep -E 'ERROR SUMMARY' gpurun_out/r4s_mem.log | tail -1)"
$M tests/test_sphr_velocity_gpu.py -k "advance" > gpurun_out/r4s_mem_sphr.log 2>&1
echo "memcheck spherical episodes: $(grep -E 'passed|failed' gpurun_out/r4s_mem_sphr.log | tail -1) | $(grep -E 'ERROR SUMMARY' gpurun_out/r4s_mem_sphr.log | tail -1)"
