#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 > gpurun_out/r3a_bench_c2_n8.log 2>&1; tail -1 gpurun_out/r3a_bench_c2_n8.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']; print('N=8 value %.3g ms %.2f e2e %.3g full %.3g'%(d['value'],d['ms_per_step'],e['value'],e['every_output_copied_back']['value']), d.get('host_affinity'))" || tail -5 gpurun_out/r3a_bench_c2_n8.log
nvidia-smi topo -m > gpurun_out/r3a_topo.txt 2>&1; head -10 gpurun_out/r3a_topo.txt | cut -c1-150
