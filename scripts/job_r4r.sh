#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -k "sphr or full" > gpurun_out/r4r_pytest.log 2>&1; tail -4 gpurun_out/r4r_pytest.log
timeout 900 python bench.py --config c5 --steps 5 --no-cpu-baseline > gpurun_out/r4r_bench_c5.log 2>&1; tail -1 gpurun_out/r4r_bench_c5.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c5', 'ms/step %.3f'%d['ms_per_step'], 'value %.4g'%d['value'], 'episode_frac %.3f'%r['episode_frac'], r['kernel_classes_ms_per_step'], (d.get('parity') or {}).get('per_zone_rel'))" || tail -3 gpurun_out/r4r_bench_c5.log
