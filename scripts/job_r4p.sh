#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "ragged or slow_faces or density_advance" > gpurun_out/r4p_pytest.log 2>&1; tail -2 gpurun_out/r4p_pytest.log
python bench.py --steps 10 --no-cpu-baseline --no-parity > gpurun_out/r4p_bench_c2.log 2>&1; tail -1 gpurun_out/r4p_bench_c2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2', 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%r['frac'], r['kernel_classes_ms_per_step'])"
python scripts/perf_edge.py 256 2>&1 | tail -4
