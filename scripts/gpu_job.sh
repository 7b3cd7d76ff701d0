#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
PERF_ONLY=periodic,1 timeout 200 python scripts/perf_episodes.py 128 2>&1 | grep advance | cut -c1-100
