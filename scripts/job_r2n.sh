#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_registry_gpu.py -x -q > gpurun_out/r2n_pytest.log 2>&1; tail -5 gpurun_out/r2n_pytest.log
python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2n_bench.log 2>&1; tail -1 gpurun_out/r2n_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(json.dumps(d['e2e'])[:1500])"
