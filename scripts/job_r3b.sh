#!/bin/bash
mkdir -p gpurun_out
for c in c3 c4 c5; do
  python bench.py --config $c --steps 5 --no-cpu-baseline > gpurun_out/r3b_bench_$c.log 2>&1
  tail -1 gpurun_out/r3b_bench_$c.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$c', 'ms/step %.2f'%d['ms_per_step'], 'value %.3g'%d['value'], 'episode_frac %.3f'%r['episode_frac'], {k:round(v,2) for k,v in r['kernel_classes_ms_per_step'].items()}, 'parity', d.get('parity',{}).get('worst_per_zone'))" || tail -5 gpurun_out/r3b_bench_$c.log
done
