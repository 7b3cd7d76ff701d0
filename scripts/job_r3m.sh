#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r3m_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3m_smoke.log
python bench.py > gpurun_out/r3m_bench_c2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r3m_bench_c2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2', 'ms/step %.3f'%d['ms_per_step'], 'value %.4g'%d['value'], 'frac %.3f'%r['frac'], 'episode_frac %.3f'%r['episode_frac'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline']['value'], 'parity', d['parity']['worst_max_norm'], d['parity']['worst_per_zone'], 'exact', d['exact_build']['ms_per_step'], d['clocks'])"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r3m_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/r3m_bench_ref.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3m_pytest.log 2>&1; tail -3 gpurun_out/r3m_pytest.log
