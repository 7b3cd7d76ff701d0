#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; tail -2 gpurun_out/r2h_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r2h_bench.log 2>&1; tail -1 gpurun_out/r2h_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value %.4g ms/step %.3f e2e %.4g roofline %.3f episode %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['episode_frac']))
print(d['roofline']['kernel_classes_ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2h_ncu_bench.log 2>&1
