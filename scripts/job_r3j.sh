#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "test_fused_edge_2d and march" > gpurun_out/r3j_pytest.log 2>&1; tail -6 gpurun_out/r3j_pytest.log
for t in 2 3; do
  python bench.py --config c4 --steps 5 --no-cpu-baseline --no-parity --opt tile2d=$t > gpurun_out/r3j_bench_c4_t$t.log 2>&1
  tail -1 gpurun_out/r3j_bench_c4_t$t.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c4 tile2d=$t', 'ms/step %.2f'%d['ms_per_step'], {k:round(v,2) for k,v in r['kernel_classes_ms_per_step'].items()})" 2>/dev/null || tail -3 gpurun_out/r3j_bench_c4_t$t.log | cut -c1-300
done
