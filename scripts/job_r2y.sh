#!/bin/bash
# compute-sanitizer memcheck + racecheck of the kernels added late in round 2 (cell-centred mkutrans / velpred with their
# shared-memory exchange, EOS kernels, average) + the multi-rank selftest + the full GPU suite
mkdir -p gpurun_out
run() {  # tool, log tag, test files, pytest -k expression
  timeout 1500 compute-sanitizer --tool $1 --print-limit 20 python -m pytest $3 -x -q -k "$4" > gpurun_out/r2y_$2.log 2>&1
  echo "== $1 $2: $(grep -E 'passed|failed' gpurun_out/r2y_$2.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2y_$2.log | tail -1)"
}
run memcheck mem_velpred tests/test_parity_gpu.py "test_mkutrans_velpred"
run racecheck race_velpred tests/test_parity_gpu.py "test_mkutrans_velpred and ppm1"
run memcheck mem_eos "tests/test_eos.py tests/test_average.py" "not exact"
run racecheck race_eos "tests/test_eos.py tests/test_average.py" "firstdt or average or make_etarho"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --selftest > gpurun_out/r2y_selftest.log 2>&1; tail -1 gpurun_out/r2y_selftest.log | cut -c1-600
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest.log 2>&1; tail -3 gpurun_out/r2y_pytest.log
