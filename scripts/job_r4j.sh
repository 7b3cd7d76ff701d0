#!/bin/bash
# final single-GPU measurements of round 2 (session 2): tests, smoke, the four bench lines, the CPU arm, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r4j_pytest.log 2>&1; tail -3 gpurun_out/r4j_pytest.log
python __graft_entry__.py smoke > gpurun_out/r4j_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r4j_smoke.log | cut -c1-300
python bench.py > gpurun_out/r4j_bench_c2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r4j_bench_c2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2', 'ms/step %.3f'%d['ms_per_step'], 'value %.4g'%d['value'], 'frac %.3f'%r['frac'], 'episode_frac %.3f'%r['episode_frac'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline']['value'], 'exact', d['exact_build']['ms_per_step'], d['clocks'])"
for c in c3 c4 c5; do
timeout 900 python bench.py --config $c --steps 5 --no-cpu-baseline > gpurun_out/r4j_bench_$c.log 2>&1; tail -1 gpurun_out/r4j_bench_$c.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$c', 'ms/step %.3f'%d['ms_per_step'], 'value %.4g'%d['value'], 'episode_frac %.3f'%r['episode_frac'], r['kernel_classes_ms_per_step'], (d.get('parity') or {}).get('per_zone_rel'))" || tail -3 gpurun_out/r4j_bench_$c.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4j_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r4j_ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_edge3 -s 1 -c 1 -o gpurun_out/prof_r4_edge3_ppm2 python scripts/one_edge.py 256 2 3 > gpurun_out/r4j_ncu_ppm2.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_r4_edge3_ppm2.ncu-rep --page raw --csv > gpurun_out/prof_r4_edge3_ppm2_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r4_edge3_ppm2.ncu-rep --page source --csv > gpurun_out/prof_r4_edge3_ppm2_src.csv 2>/dev/null
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r4j_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/r4j_bench_ref.log | cut -c1-400
