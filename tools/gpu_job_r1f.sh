#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lifting_gpu.py -x -q 2>&1 | tail -6
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 --dump gpurun_out/bench_x3.json 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --precision bf16 --dump gpurun_out/bench_bf16.json 2>&1 | tail -1 | cut -c1-300
python tools/summarize_bench.py gpurun_out/bench_x3.json > gpurun_out/summary_x3.txt; head -16 gpurun_out/summary_x3.txt
python tools/summarize_bench.py gpurun_out/bench_bf16.json > gpurun_out/summary_bf16.txt; head -16 gpurun_out/summary_bf16.txt
EGOTAP_PU=steps EGOTAP_SKIP_DUMMY=0 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-200
