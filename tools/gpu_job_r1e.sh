#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python tools/attn_only.py 256
timeout 120 python tools/attn_only.py 256 bf16
timeout 600 python bench.py --steps 20 --warmup 3 --dump gpurun_out/bench_x3.json 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --precision bf16 --dump gpurun_out/bench_bf16.json 2>&1 | tail -1 | cut -c1-300
python tools/summarize_bench.py gpurun_out/bench_x3.json | head -12
python tools/summarize_bench.py gpurun_out/bench_bf16.json | head -12
