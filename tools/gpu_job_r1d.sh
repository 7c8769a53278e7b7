#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -k fused_attention 2>&1 | tail -3
timeout 120 python tools/attn_only.py 256
timeout 120 python tools/attn_only.py 256 bf16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 3 -c 1 -o gpurun_out/attn_full python tools/attn_only.py 64 > gpurun_out/attn_full.log 2>&1
tail -2 gpurun_out/attn_full.log
