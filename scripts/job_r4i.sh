#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu -k "fused_edge or density_advance or c3 or full or slow_faces or thin_end" > gpurun_out/r4i_pytest.log 2>&1; tail -4 gpurun_out/r4i_pytest.log
timeout 600 python bench.py --config c3 --steps 5 --no-cpu-baseline > gpurun_out/r4i_bench_c3.log 2>&1; tail -1 gpurun_out/r4i_bench_c3.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c3 ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_classes_ms_per_step'], d['parity']['per_zone_rel'])" || tail -5 gpurun_out/r4i_bench_c3.log
PERF_ONLY=periodic,2 PERF_EPISODES=density timeout 300 python scripts/perf_episodes.py 256 > gpurun_out/r4i_perf256_ppm2.log 2>&1; grep advance gpurun_out/r4i_perf256_ppm2.log
