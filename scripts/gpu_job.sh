#!/bin/bash
# scratch job run on the GPU box by gpurun (edited per experiment)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/l_bench_32x8.log 2>&1
timeout 200 python bench.py --no-cpu-baseline --opt fused_by=1616 > gpurun_out/l_bench_16x16.log 2>&1
for f in gpurun_out/l_bench_32x8.log gpurun_out/l_bench_16x16.log; do python - $f <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1]); print(sys.argv[1], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['kernel_classes_ms_per_step'], d['e2e']['value'])
P
done
