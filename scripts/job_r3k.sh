#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "(test_fused_edge_2d or test_fused_edge_ragged_boxes or test_fused_edge_slow_faces) and (2- or -2)" > gpurun_out/r3k_pytest.log 2>&1; tail -3 gpurun_out/r3k_pytest.log
for c in c4 c3; do
  python bench.py --config $c --steps 5 --no-cpu-baseline > gpurun_out/r3k_bench_$c.log 2>&1
  tail -1 gpurun_out/r3k_bench_$c.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$c', 'ms/step %.2f'%d['ms_per_step'], {k:round(v,2) for k,v in r['kernel_classes_ms_per_step'].items()}, 'parity', d.get('parity',{}).get('worst_max_norm'))" 2>/dev/null || tail -3 gpurun_out/r3k_bench_$c.log | cut -c1-300
done
