#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py tests/test_sphr_velocity_gpu.py -q -m gpu -k "thin_end or ragged or c3 or full or sphr or density_advance" > gpurun_out/r4f_pytest.log 2>&1; tail -4 gpurun_out/r4f_pytest.log
for o in 1 0; do
timeout 600 python bench.py --config c3 --steps 5 --no-cpu-baseline --opt pair_streams=$o > gpurun_out/r4f_bench_c3_p$o.log 2>&1; tail -1 gpurun_out/r4f_bench_c3_p$o.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c3 pair_streams=$o ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_classes_ms_per_step'], d['parity']['per_zone_rel'])" || tail -5 gpurun_out/r4f_bench_c3_p$o.log
done
for o in 1 0; do
timeout 900 python bench.py --config c5 --steps 3 --no-cpu-baseline --no-parity --opt pair_streams=$o --opt thin_edge=$o > gpurun_out/r4f_bench_c5_p$o.log 2>&1; tail -1 gpurun_out/r4f_bench_c5_p$o.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c5 pair+thin=$o ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_classes_ms_per_step'])" || tail -5 gpurun_out/r4f_bench_c5_p$o.log
done
