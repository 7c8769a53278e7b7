#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for b in 1 32; do
timeout 600 python bench.py --steps 50 --warmup 5 --batch $b --dump gpurun_out/bench_x3_b$b.json 2>&1 | tail -1 | cut -c1-250
python tools/summarize_bench.py gpurun_out/bench_x3_b$b.json 2>/dev/null | head -14
done
