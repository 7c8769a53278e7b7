#!/bin/bash
# Round-2 job: per-kernel tables at the reference's evaluation batch sizes (16 / 32) and at batch 1.
set -x
mkdir -p gpurun_out
for b in 1 16 32; do
  timeout 300 python bench.py --batch $b --steps 50 --warmup 5 --only-headline --dump gpurun_out/r2z_b$b.json > /dev/null 2>&1
  python tools/summarize_bench.py gpurun_out/r2z_b$b.json 2>/dev/null | head -22 | cut -c1-200
done
