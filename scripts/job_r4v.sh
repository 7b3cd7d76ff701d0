#!/bin/bash
# last check of the session: the whole GPU suite, smoke, the headline line with the final library
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r4v_pytest.log 2>&1; tail -3 gpurun_out/r4v_pytest.log
python __graft_entry__.py smoke > gpurun_out/r4v_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r4v_smoke.log | cut -c1-300
python bench.py > gpurun_out/r4v_bench_c2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r4v_bench_c2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2', 'ms/step %.3f'%d['ms_per_step'], 'value %.4g'%d['value'], 'frac %.3f'%r['frac'], 'episode_frac %.3f'%r['episode_frac'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline']['value'], 'exact', d['exact_build']['ms_per_step'], d['clocks'])"
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest -x -q tests/test_parity_gpu.py -k "test_density_advance and fast and 3-" > gpurun_out/r4v_race_episode.log 2>&1
echo "race density_advance episodes (final library): $(grep -E 'passed|failed' gpurun_out/r4v_race_episode.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r4v_race_episode.log | tail -1)"
