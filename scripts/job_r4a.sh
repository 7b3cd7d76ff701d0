#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_multibox_gpu.py -q -m gpu -k "premac or mkutrans" > gpurun_out/r4a_pytest.log 2>&1; tail -4 gpurun_out/r4a_pytest.log
for o in 1 0; do
PERF_OPTS=premac_fuse=$o PERF_ONLY=periodic,1 timeout 600 python scripts/perf_episodes.py 256 > gpurun_out/r4a_perf256_f$o.log 2>&1; grep -i "premac\|velocity" gpurun_out/r4a_perf256_f$o.log
done
PERF_ONLY=periodic,2 timeout 600 python scripts/perf_episodes.py 256 > gpurun_out/r4a_perf256_ppm2.log 2>&1; grep -i "premac\|velocity" gpurun_out/r4a_perf256_ppm2.log
PERF_ONLY=periodic,1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r4a_launches_premac.csv python scripts/perf_episodes.py 256 > gpurun_out/r4a_ncu.log 2>&1; tail -2 gpurun_out/r4a_ncu.log
