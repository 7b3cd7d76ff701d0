#!/bin/bash
# compute-sanitizer on the kernels / paths added at the very end of round 2
mkdir -p gpurun_out
run() {  # tool, log tag, test files, pytest -k expression
  timeout 1500 compute-sanitizer --tool $1 --print-limit 20 python -m pytest $3 -x -q -k "$4" > gpurun_out/r3r_$2.log 2>&1
  echo "== $1 $2: $(grep -E 'passed|failed' gpurun_out/r3r_$2.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r3r_$2.log | tail -1)"
}
run racecheck race_march tests/test_parity_gpu.py "test_fused_edge_2d and march and (shape0 or shape3) and 2-"
run memcheck mem_march tests/test_parity_gpu.py "test_fused_edge_2d and march"
run memcheck mem_multibox tests/test_multibox_gpu.py "several_boxes"
run memcheck mem_bds tests/test_parity_gpu.py "test_bds"
run racecheck race_vpfast tests/test_parity_gpu.py "test_mkutrans_velpred and fast and inout2"
