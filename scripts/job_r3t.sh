#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_registry_gpu.py tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -k "registry or density_advance or full or e2e or resid" > gpurun_out/r3t_pytest.log 2>&1; tail -4 gpurun_out/r3t_pytest.log
for o in 1 0; do
python bench.py --steps 5 --no-cpu-baseline --no-parity --e2e-steps 4 --opt async_upload=$o > gpurun_out/r3t_bench_a$o.log 2>&1
tail -1 gpurun_out/r3t_bench_a$o.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('async_upload=$o', 'ms/step %.3f'%d['ms_per_step'], 'e2e %.4g'%e['value'], 'full %.4g'%e['every_output_copied_back']['value'], 'checksum', e['check_sum_rho_new'], e['every_output_copied_back']['check_sum_rho_new'])" || tail -5 gpurun_out/r3t_bench_a$o.log
done
