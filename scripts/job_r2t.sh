#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2t_smoke.log
python bench.py > gpurun_out/r2t_bench_c2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2t_bench_c2.log | cut -c1-1500
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2t_launches_c5.csv python bench.py --config c5 --n 256 --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2t_ncu_c5.log 2>&1; echo "ncu rc=$?"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; tail -2 gpurun_out/r2t_pytest.log
