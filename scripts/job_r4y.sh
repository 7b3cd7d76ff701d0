#!/bin/bash
# 2 GPUs, final library: multi-rank selftest (now with the FAST spherical case) and the c5 line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $TR bench.py --gpus 2 --selftest > gpurun_out/r4y_selftest_n2.log 2>&1; tail -1 gpurun_out/r4y_selftest_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('selftest ok', d['ok'], len(d['cases']), [k for k,v in d['cases'].items() if v['rc']!=0]); print(d['cases'].get('dm3_sphr_ppm1_fast'))" || tail -3 gpurun_out/r4y_selftest_n2.log | cut -c1-300
timeout 600 $TR bench.py --gpus 2 --config c5 --steps 5 --no-cpu-baseline --no-parity > gpurun_out/r4y_bench_c5_n2.log 2>&1
tail -1 gpurun_out/r4y_bench_c5_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c5 N=2', 'ms/step %.2f'%d['ms_per_step'], 'value %.4g'%d['value'], {k:round(v,2) for k,v in r['kernel_classes_ms_per_step'].items()})" || tail -3 gpurun_out/r4y_bench_c5_n2.log | cut -c1-300
