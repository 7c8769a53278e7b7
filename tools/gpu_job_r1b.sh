#!/bin/bash
# round-1 GPU job B: fused attention bring-up (own subprocess + timeout), then full tests and bench
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -k fused_attention 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 --dump gpurun_out/bench_x3.json 2>&1 | tail -1 | cut -c1-1800
timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --dump gpurun_out/bench_bf16.json 2>&1 | tail -1 | cut -c1-1800
EGOTAP_ATTN=unfused timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
