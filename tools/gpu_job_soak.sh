#!/bin/bash
set -x
mkdir -p gpurun_out
for prec in bf16x3 bf16; do for B in 8 64 256; do timeout 300 python tools/attn_check.py $B $prec 400 2>&1 | tail -1 | cut -c1-200; done; done
timeout 1200 python tools/soak.py 150 2>&1 | tail -20
