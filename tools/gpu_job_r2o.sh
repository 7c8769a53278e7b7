#!/bin/bash
# Round-2 job: attention kernel with two MMA-issuing warps + per-role register budgets against the committed single-issuer kernel
# (tools/libegotap_b200_base.so = HEAD), attention tests, one ncu capture, headline bench in both precisions.
#   gpurun --timeout 1500 -- 'bash tools/gpu_job_r2o.sh > gpurun_out/r2o.log 2>&1'
set -x
mkdir -p gpurun_out
for rep in 1 2; do
for prec in bf16x3 bf16; do
  EGOTAP_B200_LIB=$PWD/tools/libegotap_b200_base.so timeout 300 python tools/attn_only.py 256 $prec 2>&1 | tail -1 | sed 's/^/base  /'
  timeout 300 python tools/attn_only.py 256 $prec 2>&1 | tail -1 | sed 's/^/new   /'
done; done
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_lifting_gpu.py -m gpu -q -x -rfEs 2>&1 | tail -5
for prec in bf16x3 bf16; do
  timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --only-headline --dump gpurun_out/r2o_bench_${prec}.json 2>&1 | tail -1 | cut -c1-200
  python tools/summarize_bench.py gpurun_out/r2o_bench_${prec}.json 2>/dev/null | head -8
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -c 1 \
    -o gpurun_out/r2o_attention python tools/attn_only.py 64 bf16x3 > gpurun_out/r2o_ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -c 1 \
    -o gpurun_out/r2o_attention_bf16 python tools/attn_only.py 64 bf16 > gpurun_out/r2o_ncu_attn_bf16.log 2>&1
tail -2 gpurun_out/r2o_ncu_attn.log | cut -c1-200
