#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; tail -2 gpurun_out/r2l_pytest.log
for c in c3 c4; do python bench.py --config $c --steps 5 --no-cpu-baseline > gpurun_out/r2l_bench_$c.log 2>&1; tail -1 gpurun_out/r2l_bench_$c.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$c', 'ms/step %.2f'%d['ms_per_step'], 'value %.3g'%d['value'], {k:round(v,3) for k,v in r['kernel_classes_ms_per_step'].items()}, d.get('parity',{}).get('worst'))"; done
python bench.py --config c5 --steps 3 --no-cpu-baseline > gpurun_out/r2l_bench_c5.log 2>&1; tail -1 gpurun_out/r2l_bench_c5.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c5', 'ms/step %.2f'%d['ms_per_step'], 'value %.3g'%d['value'], {k:round(v,3) for k,v in r['kernel_classes_ms_per_step'].items()})"
