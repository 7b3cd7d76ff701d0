#!/bin/bash
# N GPUs (N = $1): the strong-scaling lines of c4 and c5
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518"
for c in c4 c5; do
  timeout 600 $TR bench.py --gpus $N --config $c --steps 5 --no-cpu-baseline --no-parity > gpurun_out/r4n_bench_${c}_n$N.log 2>&1
  tail -1 gpurun_out/r4n_bench_${c}_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$c N=$N', 'ms/step %.2f'%d['ms_per_step'], 'value %.4g'%d['value'], d['scaling'], {k:round(v,2) for k,v in r['kernel_classes_ms_per_step'].items()})" 2>/dev/null || tail -3 gpurun_out/r4n_bench_${c}_n$N.log | cut -c1-300
done
nvidia-smi topo -m > gpurun_out/r4n_topo_n$N.txt 2>&1
