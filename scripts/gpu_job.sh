#!/bin/bash
# scratch job run on the GPU box by gpurun (edited per experiment)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
