#!/bin/bash
mkdir -p gpurun_out
{
echo "== toolchain"; which gfortran flang nvfortran f2c 2>&1; ls /usr/bin/*fortran* 2>&1 | head
echo "== topo"; nvidia-smi topo -m 2>&1 | head -30
echo "== numa"; (numactl --hardware 2>&1 || lscpu | grep -i -E "numa|^CPU\(s\)|Model name|Socket") | head -30
echo "== nproc"; nproc; free -g | head -3
} > gpurun_out/r2a_env.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; tail -3 gpurun_out/r2a_pytest.log
python bench.py > gpurun_out/r2a_bench.log 2>&1; tail -1 gpurun_out/r2a_bench.log | cut -c1-600
