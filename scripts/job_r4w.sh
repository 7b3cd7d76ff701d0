#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu -k "density_advance or flux_update or full" > gpurun_out/r4w_pytest.log 2>&1; tail -2 gpurun_out/r4w_pytest.log
python bench.py --steps 10 --no-cpu-baseline --no-parity > gpurun_out/r4w_bench_c2.log 2>&1; tail -1 gpurun_out/r4w_bench_c2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2', 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%r['frac'], 'episode_frac %.3f'%r['episode_frac'], r['kernel_classes_ms_per_step'], 'e2e %.4g'%d['e2e']['value'])"
