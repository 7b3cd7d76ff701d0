#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py tests/test_registry_gpu.py -q -m gpu -k "fused_edge or density_advance or full or slow_faces or resid or registry or e2e" > gpurun_out/r4o_pytest.log 2>&1; tail -4 gpurun_out/r4o_pytest.log
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r4o_bench_c2.log 2>&1; tail -1 gpurun_out/r4o_bench_c2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2', 'ms/step %.3f'%d['ms_per_step'], 'value %.4g'%d['value'], 'frac %.3f'%r['frac'], 'episode_frac %.3f'%r['episode_frac'], r['kernel_classes_ms_per_step'], d['parity']['per_zone_rel'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_edge3 -s 1 -c 1 -o gpurun_out/prof_r4_edge3_ppm1 python scripts/one_edge.py 256 1 3 > gpurun_out/r4o_ncu_ppm1.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_r4_edge3_ppm1.ncu-rep --page raw --csv > gpurun_out/prof_r4_edge3_ppm1_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r4_edge3_ppm1.ncu-rep --page source --csv > gpurun_out/prof_r4_edge3_ppm1_src.csv 2>/dev/null
