#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_parity_gpu.py -x -q -k "test_mkutrans_velpred and (periodic or inout2) and n1" > gpurun_out/r2z_race_velpred.log 2>&1
echo "== racecheck velpred: $(grep -E 'passed|failed' gpurun_out/r2z_race_velpred.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r2z_race_velpred.log | tail -1)"
nvidia-smi topo -m > gpurun_out/r2z_topo.txt 2>&1; head -12 gpurun_out/r2z_topo.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2z_bench_c2_n2.log 2>&1; tail -1 gpurun_out/r2z_bench_c2_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']; print('N=2 value %.3g ms %.2f e2e %.3g h2d %.3g d2h %.3g full %.3g'%(d['value'],d['ms_per_step'],e['value'],e['h2d_bytes_per_step'],e['d2h_bytes_per_step'],e['every_output_copied_back']['value']))" || tail -5 gpurun_out/r2z_bench_c2_n2.log
