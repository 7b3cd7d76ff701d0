#!/bin/bash
mkdir -p gpurun_out
python bench.py --bds --steps 3 --no-cpu-baseline --no-parity > gpurun_out/r3q_bench.log 2>&1
tail -1 gpurun_out/r3q_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('bds unrolled', 'ms/step %.2f'%d['ms_per_step'], {k:round(v,2) for k,v in r['kernel_classes_ms_per_step'].items()})"
