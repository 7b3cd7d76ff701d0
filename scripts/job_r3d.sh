#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_bds_conc -s 3 -c 1 -o gpurun_out/prof_r02_bds_conc python bench.py --bds --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r3d_ncu1.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_bds_slope -s 1 -c 1 -o gpurun_out/prof_r02_bds_slope python bench.py --bds --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r3d_ncu2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_r02_bds*
