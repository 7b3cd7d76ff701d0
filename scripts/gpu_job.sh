#!/bin/bash
(time python -m pytest tests/test_full_size_gpu.py -x -q -k c3 2>&1 | tail -12) 2>&1
