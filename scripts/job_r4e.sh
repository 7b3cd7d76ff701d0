#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu -k "thin_end or ragged or c3 or full" > gpurun_out/r4e_pytest.log 2>&1; tail -4 gpurun_out/r4e_pytest.log
for o in 1 0; do
timeout 600 python bench.py --config c3 --steps 5 --no-cpu-baseline --opt thin_edge=$o > gpurun_out/r4e_bench_c3_t$o.log 2>&1; tail -1 gpurun_out/r4e_bench_c3_t$o.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('thin_edge=$o ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_classes_ms_per_step'], d.get('parity'))" || tail -5 gpurun_out/r4e_bench_c3_t$o.log
done
