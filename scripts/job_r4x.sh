#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r4x_bench_c2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r4x_bench_c2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2', 'ms/step %.3f'%d['ms_per_step'], 'value %.4g'%d['value'], 'frac %.3f'%r['frac'], 'episode_frac %.3f'%r['episode_frac'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline']['value'], 'exact', d['exact_build']['ms_per_step'], r['other_kernels']['update']['frac'], d['clocks'])"
timeout 600 python bench.py --config c3 --steps 5 --no-cpu-baseline > gpurun_out/r4x_bench_c3.log 2>&1; tail -1 gpurun_out/r4x_bench_c3.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c3', 'ms/step %.3f'%d['ms_per_step'], 'value %.4g'%d['value'], 'episode_frac %.3f'%r['episode_frac'], r['kernel_classes_ms_per_step'])"
