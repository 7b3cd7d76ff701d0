#!/bin/bash
# final check of the second session: the whole GPU suite, smoke, the headline line and c5 after the last changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r4q_pytest.log 2>&1; tail -3 gpurun_out/r4q_pytest.log
python __graft_entry__.py smoke > gpurun_out/r4q_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r4q_smoke.log | cut -c1-300
python bench.py > gpurun_out/r4q_bench_c2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r4q_bench_c2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2', 'ms/step %.3f'%d['ms_per_step'], 'value %.4g'%d['value'], 'frac %.3f'%r['frac'], 'episode_frac %.3f'%r['episode_frac'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline']['value'], 'exact', d['exact_build']['ms_per_step'], d['clocks'])"
timeout 900 python bench.py --config c5 --steps 5 --no-cpu-baseline > gpurun_out/r4q_bench_c5.log 2>&1; tail -1 gpurun_out/r4q_bench_c5.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c5', 'ms/step %.3f'%d['ms_per_step'], 'value %.4g'%d['value'], 'kernel', r['kernel'], 'frac %.3f'%r['frac'], 'episode_frac %.3f'%r['episode_frac'], r['kernel_classes_ms_per_step'])" || tail -3 gpurun_out/r4q_bench_c5.log
