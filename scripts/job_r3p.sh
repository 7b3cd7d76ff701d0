#!/bin/bash
mkdir -p gpurun_out
for c in c4 c5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --config $c --steps 5 --no-cpu-baseline > gpurun_out/r3p_bench_${c}_n2.log 2>&1
  tail -1 gpurun_out/r3p_bench_${c}_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$c N=2', 'ms/step %.2f'%d['ms_per_step'], 'value %.3g'%d['value'], d['scaling'], {k:round(v,2) for k,v in r['kernel_classes_ms_per_step'].items()})" 2>/dev/null || tail -3 gpurun_out/r3p_bench_${c}_n2.log | cut -c1-300
done
