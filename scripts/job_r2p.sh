#!/bin/bash
# final captures of round 2: ncu --set full of the edge kernel variants and the update kernel, launch list of the bench
# episode, racecheck of the episodes and the 2-D kernel, bench lines
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_fused_edge3 -s 1 -c 1 -o gpurun_out/prof_r02_edge3_ppm1 python scripts/one_edge.py 256 1 3 > gpurun_out/r2p_ncu1.log 2>&1
$NCU -k regex:k_fused_edge3 -s 1 -c 1 -o gpurun_out/prof_r02_edge3_ppm2 python scripts/one_edge.py 256 2 3 > gpurun_out/r2p_ncu2.log 2>&1
$NCU -k regex:k_fused_edge3 -s 6 -c 1 -o gpurun_out/prof_r02_edge3_xf1 python scripts/one_episode.py 256 2 > gpurun_out/r2p_ncu3.log 2>&1
$NCU -k regex:k_flux_update3 -s 1 -c 1 -o gpurun_out/prof_r02_fluxupd python scripts/one_episode.py 256 2 > gpurun_out/r2p_ncu4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2p_ncu_bench.log 2>&1
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_parity_gpu.py -x -q -k "test_density_advance and fast and 3-" > gpurun_out/r2p_race_episode.log 2>&1
echo "race episode: $(grep -E 'passed|failed' gpurun_out/r2p_race_episode.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r2p_race_episode.log | tail -1)"
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_parity_gpu.py -x -q -k "2d and fast" > gpurun_out/r2p_race_2d.log 2>&1
echo "race 2d: $(grep -E 'passed|failed' gpurun_out/r2p_race_2d.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r2p_race_2d.log | tail -1)"
python bench.py > gpurun_out/r2p_bench_c2.log 2>&1; tail -1 gpurun_out/r2p_bench_c2.log | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2p_bench_ref.log 2>&1; tail -1 gpurun_out/r2p_bench_ref.log | cut -c1-300
