#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r2k_bench_c2.log 2>&1; tail -1 gpurun_out/r2k_bench_c2.log | cut -c1-1500
for c in c3 c4; do python bench.py --config $c --steps 5 > gpurun_out/r2k_bench_$c.log 2>&1; tail -1 gpurun_out/r2k_bench_$c.log | cut -c1-900; done
python bench.py --config c5 --steps 3 --no-cpu-baseline > gpurun_out/r2k_bench_c5.log 2>&1; tail -2 gpurun_out/r2k_bench_c5.log | cut -c1-900
