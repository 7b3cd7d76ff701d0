#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu -k "2d or 2-" > gpurun_out/r4h_pytest.log 2>&1; tail -4 gpurun_out/r4h_pytest.log
timeout 600 python bench.py --config c4 --steps 5 --no-cpu-baseline > gpurun_out/r4h_bench_c4.log 2>&1; tail -1 gpurun_out/r4h_bench_c4.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c4 ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_classes_ms_per_step'], d['parity']['per_zone_rel'])" || tail -5 gpurun_out/r4h_bench_c4.log
