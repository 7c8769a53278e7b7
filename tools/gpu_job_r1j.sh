#!/bin/bash
set -x
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -k fused_attention 2>&1 | tail -5
timeout 120 python tools/attn_only.py 256
timeout 120 python tools/attn_only.py 256 bf16
timeout 120 python tools/attn_trace.py
