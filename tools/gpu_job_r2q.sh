#!/bin/bash
# Round-2 job: two-issuer attention with per-warp kv_full barriers (bit-compare with the single-issuer build), leaner GELU,
# L2 prefetch of the residual rows; GPU suite; headline bench in both precisions; ncu of the attention kernel.
set -x
mkdir -p gpurun_out
for prec in bf16x3 bf16; do
  for B in 8 64 256; do
    EGOTAP_B200_LIB=$PWD/tools/libegotap_b200_base.so timeout 120 python tools/attn_check.py $B $prec 10 2>&1 | tail -1 | cut -c1-300 | sed "s/^/base /"
    timeout 120 python tools/attn_check.py $B $prec 30 2>&1 | tail -1 | cut -c1-300 | sed "s/^/cur  /"
  done
done
for rep in 1 2; do
for prec in bf16x3 bf16; do
  EGOTAP_B200_LIB=$PWD/tools/libegotap_b200_base.so timeout 300 python tools/attn_only.py 256 $prec 2>&1 | tail -1 | sed 's/^/base  /'
  timeout 300 python tools/attn_only.py 256 $prec 2>&1 | tail -1 | sed 's/^/new   /'
done; done
timeout 900 python -m pytest tests -m gpu -q -x -rfEs 2>&1 | tail -8
for prec in bf16x3 bf16; do
  EGOTAP_B200_LIB=$PWD/tools/libegotap_b200_base.so timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --only-headline --dump gpurun_out/r2q_bench_${prec}_base.json 2>&1 | tail -1 | cut -c1-200
  python tools/summarize_bench.py gpurun_out/r2q_bench_${prec}_base.json 2>/dev/null | head -16
  timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --only-headline --dump gpurun_out/r2q_bench_${prec}.json 2>&1 | tail -1 | cut -c1-200
  python tools/summarize_bench.py gpurun_out/r2q_bench_${prec}.json 2>/dev/null | head -16
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -c 1 \
    -o gpurun_out/r2q_attention python tools/attn_only.py 64 bf16x3 > gpurun_out/r2q_ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -c 1 \
    -o gpurun_out/r2q_attention_bf16 python tools/attn_only.py 64 bf16 > gpurun_out/r2q_ncu_attn_bf16.log 2>&1
tail -2 gpurun_out/r2q_ncu_attn.log | cut -c1-200
