#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "bds" > gpurun_out/r3h_pytest.log 2>&1; tail -2 gpurun_out/r3h_pytest.log
for blk in 32,2,2 32,4,1 32,1,4 64,2,1; do
  MGPU_BDS_BLOCK=$blk python bench.py --bds --steps 3 --no-cpu-baseline --no-parity > gpurun_out/r3h_bench_$blk.log 2>&1
  tail -1 gpurun_out/r3h_bench_$blk.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('block $blk', 'ms/step %.2f'%d['ms_per_step'], {k:round(v,2) for k,v in r['kernel_classes_ms_per_step'].items()})" 2>/dev/null || tail -2 gpurun_out/r3h_bench_$blk.log | cut -c1-300
done
