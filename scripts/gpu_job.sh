#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for o in periodic,1 walls,2; do PERF_ONLY=$o timeout 200 python scripts/perf_episodes.py 128 2>&1 | grep advance | cut -c1-100; done
