#!/bin/bash
# compute-sanitizer memcheck + racecheck of the fused edge kernels (every template family) at small sizes
mkdir -p gpurun_out
run() {  # tool, log tag, pytest -k expression
  timeout 1500 compute-sanitizer --tool $1 --print-limit 20 python -m pytest tests/test_parity_gpu.py -x -q -k "$3" > gpurun_out/r2o_$2.log 2>&1
  echo "== $1 $2: $(grep -E 'passed|failed' gpurun_out/r2o_$2.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2o_$2.log | tail -1)"
}
run memcheck mem_edge3d "test_fused_edge_ragged_boxes and fast"
run racecheck race_edge3d "test_fused_edge_ragged_boxes and fast and (shape1 or shape2)"
run memcheck mem_episode "test_density_advance and fast and 3-"
run racecheck race_episode "test_density_advance and fast and 3- and ppm1"
