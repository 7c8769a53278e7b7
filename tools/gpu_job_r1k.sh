#!/bin/bash
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --workload e2e_rgb --batch 256 --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-1500
