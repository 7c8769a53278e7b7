#!/bin/bash
# Round-2 job 3: two-group attention softmax, GEMM epilogue fixes (no GPU-scope fence on the accumulator hand-back, scale / bias
# staged in shared memory, residual one chunk ahead), tall reduce_partials.
#   gpurun --timeout 2400 -- 'bash tools/gpu_job_r2e.sh > gpurun_out/r2e.log 2>&1'
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rfEs 2>&1 | tail -30 > gpurun_out/r2e_pytest.log; tail -8 gpurun_out/r2e_pytest.log
for prec in bf16x3 bf16; do timeout 300 python tools/attn_only.py 256 $prec 2>&1 | tail -1; done
for prec in bf16x3 bf16; do
  timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --only-headline --dump gpurun_out/r2e_bench_${prec}.json 2>&1 | tail -1 | cut -c1-200
  python tools/summarize_bench.py gpurun_out/r2e_bench_${prec}.json 2>/dev/null | head -16
done
for epi in rows coalesced; do
  EGOTAP_EPI=$epi timeout 600 python bench.py --precision bf16 --steps 20 --warmup 3 --only-headline --dump gpurun_out/r2e_bench_bf16_$epi.json 2>&1 | tail -1 | cut -c1-200
  python tools/summarize_bench.py gpurun_out/r2e_bench_bf16_$epi.json 2>/dev/null | head -12
done
timeout 600 python bench.py --workload train --precision bf16 --batch 256 --steps 10 --warmup 3 --dump gpurun_out/r2e_train_bf16_b256.json 2>&1 | tail -1 | cut -c1-1800
timeout 600 python bench.py --workload train --precision bf16 --batch 32 --steps 10 --warmup 3 --dump gpurun_out/r2e_train_bf16_b32.json 2>&1 | tail -1 | cut -c1-1800
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -c 1 \
    -o gpurun_out/r2e_attention python tools/attn_only.py 64 bf16x3 > gpurun_out/r2e_ncu_attn.log 2>&1
tail -2 gpurun_out/r2e_ncu_attn.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -c 1 \
    -o gpurun_out/r2e_attention_bf16 python tools/attn_only.py 64 bf16 > gpurun_out/r2e_ncu_attn_bf16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:gemm_tc --launch-skip 1 -c 4 \
    -o gpurun_out/r2e_gemm_bf16 python tools/one_step.py --precision bf16 > gpurun_out/r2e_ncu_gemm.log 2>&1
tail -2 gpurun_out/r2e_ncu_gemm.log | cut -c1-200
