#!/bin/bash
mkdir -p gpurun_out
python bench.py --config c5 --steps 3 --no-cpu-baseline > gpurun_out/r2s_bench_c5.log 2>&1; tail -1 gpurun_out/r2s_bench_c5.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c5', 'ms/step %.2f'%d['ms_per_step'], 'value %.3g'%d['value'], 'episode_frac %.3f'%r['episode_frac'], {k:round(v,3) for k,v in r['kernel_classes_ms_per_step'].items()})" || tail -5 gpurun_out/r2s_bench_c5.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; tail -2 gpurun_out/r2s_pytest.log
