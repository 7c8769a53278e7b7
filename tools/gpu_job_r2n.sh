#!/bin/bash
# Round-2 job: attention kernel with the two-step P hand-over (p_half) against the committed kernel (tools/libegotap_b200_base.so),
# GPU suite, default bench line.
#   gpurun --timeout 1500 -- 'bash tools/gpu_job_r2n.sh > gpurun_out/r2n.log 2>&1'
set -x
mkdir -p gpurun_out
for rep in 1 2; do
for prec in bf16x3 bf16; do
  EGOTAP_B200_LIB=$PWD/tools/libegotap_b200_base.so timeout 300 python tools/attn_only.py 256 $prec 2>&1 | tail -1 | sed 's/^/base  /'
  timeout 300 python tools/attn_only.py 256 $prec 2>&1 | tail -1 | sed 's/^/phalf /'
done; done
timeout 900 python -m pytest tests -m gpu -q -x -rfEs 2>&1 | tail -30 > gpurun_out/r2n_pytest.log; tail -8 gpurun_out/r2n_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --dump gpurun_out/r2n_bench_default.json 2>&1 | tail -1 | cut -c1-600
