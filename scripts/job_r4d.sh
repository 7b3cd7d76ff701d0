#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "minmax" > gpurun_out/r4d_pytest.log 2>&1; tail -2 gpurun_out/r4d_pytest.log
for c in c3 c4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r4d_launches_$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r4d_ncu_$c.log 2>&1; tail -1 gpurun_out/r4d_ncu_$c.log | cut -c1-300
done
