#!/bin/bash
(time python -m pytest tests/test_full_size_gpu.py -x -q 2>&1 | tail -8) 2>&1
