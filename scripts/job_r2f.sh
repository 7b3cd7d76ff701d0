#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; tail -1 gpurun_out/r2f_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r2f_bench.log 2>&1; tail -1 gpurun_out/r2f_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value %.4g ms/step %.3f e2e %.4g roofline %.3f episode %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['episode_frac']))
print(d['roofline']['kernel_classes_ms_per_step'])"
