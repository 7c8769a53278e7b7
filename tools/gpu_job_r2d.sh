#!/bin/bash
# Round-2 job 2: strict suite after the default flips (wide attention, mixed epilogue policy, split-K default), the new
# driver line (headline + modes.bf16 + config3), ncu --set full of the four bf16-mode layer GEMMs, launch lists with DRAM bytes.
#   gpurun --timeout 2400 -- 'bash tools/gpu_job_r2d.sh > gpurun_out/r2d.log 2>&1'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -m gpu -q -rfEs 2>&1 | tail -40 > gpurun_out/r2d_pytest.log; tail -15 gpurun_out/r2d_pytest.log
cat gpurun_out/parity.jsonl | cut -c1-400
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err; tail -3 gpurun_out/r2d_bench_default.err; cut -c1-300 gpurun_out/r2d_bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
for prec in bf16x3 bf16; do
  timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --only-headline --dump gpurun_out/r2d_bench_${prec}.json 2>&1 | tail -1 | cut -c1-200
  python tools/summarize_bench.py gpurun_out/r2d_bench_${prec}.json 2>/dev/null | head -24
done
timeout 600 python bench.py --workload e2e_rgb --batch 512 --steps 10 --warmup 3 --dump gpurun_out/r2d_e2e_rgb.json 2>&1 | tail -1 | cut -c1-600
# ncu: the four GEMMs of ViT layer 0 in bf16 mode (QKV, out-projection, MLP-up + GELU, MLP-down) of the second step
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:gemm_tc --launch-skip 1 -c 4 \
    -o gpurun_out/r2d_gemm_bf16 python tools/one_step.py --precision bf16 > gpurun_out/r2d_ncu_gemm.log 2>&1
tail -2 gpurun_out/r2d_ncu_gemm.log | cut -c1-200
for prec in bf16x3 bf16; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none --nvtx --nvtx-include "timed/" -c 60 --csv --log-file gpurun_out/r2d_launches_step_${prec}.csv \
      python tools/one_step.py --precision $prec > gpurun_out/r2d_ncu_step_${prec}.log 2>&1
  tail -1 gpurun_out/r2d_ncu_step_${prec}.log | cut -c1-200
done
