#!/bin/bash
mkdir -p gpurun_out
python bench.py --bds --steps 5 > gpurun_out/r3g_bench_c2_bds.log 2>&1
tail -1 gpurun_out/r3g_bench_c2_bds.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('bds', 'ms/step %.2f'%d['ms_per_step'], 'value %.3g'%d['value'], {k:round(v,2) for k,v in r['kernel_classes_ms_per_step'].items()}, 'parity', d.get('parity',{}).get('worst_max_norm'), 'exact', d.get('exact_build',{}).get('ms_per_step'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'e2e %.3g'%d['e2e']['value'])" || tail -5 gpurun_out/r3g_bench_c2_bds.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r3g_launches_bds.csv python bench.py --bds --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r3g_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -k "bds or velocity_advance or enthalpy_advance" > gpurun_out/r3g_pytest.log 2>&1; tail -2 gpurun_out/r3g_pytest.log
