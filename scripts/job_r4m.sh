#!/bin/bash
# N GPUs (N = $1): multi-rank parity selftest, then the bench lines of c2 (weak), c4 and c5 (strong)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --selftest > gpurun_out/r4m_selftest_n$N.log 2>&1; tail -1 gpurun_out/r4m_selftest_n$N.log | cut -c1-900
for c in c2 c4 c5; do
  timeout 900 $TR bench.py --gpus $N --config $c --steps 5 --no-cpu-baseline > gpurun_out/r4m_bench_${c}_n$N.log 2>&1
  tail -1 gpurun_out/r4m_bench_${c}_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$c N=$N', 'ms/step %.2f'%d['ms_per_step'], 'value %.4g'%d['value'], d['scaling'], 'e2e', (d.get('e2e') or {}).get('value'), {k:round(v,2) for k,v in r['kernel_classes_ms_per_step'].items()})" 2>/dev/null || tail -3 gpurun_out/r4m_bench_${c}_n$N.log | cut -c1-300
done
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests -q -m gpu -k "multi_gpu" > gpurun_out/r4m_pytest_multi.log 2>&1; tail -2 gpurun_out/r4m_pytest_multi.log; fi
