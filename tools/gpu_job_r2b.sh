#!/bin/bash
# Round-2 inference A/B job: first hardware run of the opt-in variants written after round 1's GPU budget was spent.
#   gpurun --timeout 2700 -- 'bash tools/gpu_job_r2b.sh > gpurun_out/r2b.log 2>&1'
# 1. parity of the wide attention kernel (EGOTAP_ATTN=wide) op-level and through the whole path (subprocess per case)
# 2. the attention op alone, v1 vs wide, both precisions (CUDA events)           -> is 128-key tiling faster?
# 3. bench lines of the default workload with / without it, both precisions       -> step-level effect
# 4. small-batch latency: batch 1 / 8 / 32 with and without EGOTAP_SPLITK=1, launched and replayed from a CUDA graph (--graph)
# 5. one ncu --set full capture of the wide attention kernel (tensor-pipe active %, issue stalls of the MMA warp)
# 6. the coalesced GEMM epilogue (EGOTAP_EPI=coalesced): parity, then bench lines with it alone and with both switches
set -x
export EGOTAP_STRICT_UNVERIFIED=1      # the quarantined first-hardware-run tests (tests/conftest.py) run as ordinary tests here
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 900 python -m pytest tests/test_zzz_attention_wide_gpu.py -m gpu -q 2>&1 | tail -15
cat gpurun_out/attention_wide.jsonl
timeout 1500 python -m pytest tests/test_zzz_epilogue_coalesced_gpu.py -m gpu -q 2>&1 | tail -15
cat gpurun_out/epilogue_coalesced.jsonl
for prec in bf16x3 bf16; do
  for v in v1 wide; do
    if [ $v = wide ]; then export EGOTAP_ATTN=wide; else unset EGOTAP_ATTN; fi
    echo "variant $v"; timeout 300 python tools/attn_only.py 256 $prec 2>&1 | tail -1
  done
done
unset EGOTAP_ATTN
# wait-cycle accounting of both attention kernels (a -DEB_ATTN_TRACE build of the library, tools/libtrace.so)
timeout 200 python tools/attn_trace.py 2>&1 | tail -12
EGOTAP_ATTN=wide timeout 200 python tools/attn_trace.py 2>&1 | tail -12
for prec in bf16x3 bf16; do
  timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --dump gpurun_out/r2b_bench_${prec}_v1.json 2>&1 | tail -1 | cut -c1-700
  EGOTAP_ATTN=wide timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --dump gpurun_out/r2b_bench_${prec}_wide.json 2>&1 | tail -1 | cut -c1-700
  EGOTAP_EPI=coalesced timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --dump gpurun_out/r2b_bench_${prec}_coal.json 2>&1 | tail -1 | cut -c1-700
  EGOTAP_EPI=coalesced EGOTAP_ATTN=wide timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --dump gpurun_out/r2b_bench_${prec}_both.json 2>&1 | tail -1 | cut -c1-700
done
for f in gpurun_out/r2b_bench_*.json; do echo $f; python tools/summarize_bench.py $f | head -12; done
timeout 600 python -m pytest tests/test_zzz_graph_inference_gpu.py -m gpu -q 2>&1 | tail -8
for b in 1 8 32; do
  timeout 300 python bench.py --batch $b --steps 50 --warmup 5 2>&1 | tail -1 | cut -c1-260
  EGOTAP_SPLITK=1 timeout 300 python bench.py --batch $b --steps 50 --warmup 5 2>&1 | tail -1 | cut -c1-260
  timeout 300 python bench.py --batch $b --steps 50 --warmup 5 --graph 2>&1 | tail -1 | cut -c1-260
  EGOTAP_SPLITK=1 timeout 300 python bench.py --batch $b --steps 50 --warmup 5 --graph 2>&1 | tail -1 | cut -c1-260
done
EGOTAP_ATTN=wide timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_wide -c 1 \
    -o gpurun_out/r2b_attention_wide python tools/attn_only.py 64 bf16x3 > gpurun_out/r2b_ncu_attn.log 2>&1
tail -3 gpurun_out/r2b_ncu_attn.log | cut -c1-300
