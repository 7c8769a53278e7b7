#!/bin/bash
# Round-2 final evidence job (one GPU): strict suite, the driver's default bench line, per-kernel tables, ncu launch lists with
# DRAM bytes (both precisions), ncu --set full of the four layer-0 GEMMs (bf16 mode) and of the attention kernel, small batches,
# training step, config-4 pipeline.
#   gpurun --timeout 2400 -- 'bash tools/gpu_job_r2t.sh > gpurun_out/r2t.log 2>&1'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -m gpu -q -rfEs 2>&1 | tail -40 > gpurun_out/r2t_pytest.log; tail -6 gpurun_out/r2t_pytest.log
cut -c1-400 gpurun_out/parity.jsonl | head -12
timeout 900 python bench.py > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err; tail -3 gpurun_out/r2t_bench_default.err; cut -c1-300 gpurun_out/r2t_bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
for prec in bf16x3 bf16; do
  timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --only-headline --dump gpurun_out/r2t_bench_${prec}.json 2>&1 | tail -1 | cut -c1-200
  python tools/summarize_bench.py gpurun_out/r2t_bench_${prec}.json > gpurun_out/r2t_kernel_table_${prec}.txt 2>/dev/null; head -24 gpurun_out/r2t_kernel_table_${prec}.txt
done
for prec in bf16x3 bf16; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none --nvtx --nvtx-include "timed/" -c 60 --csv --log-file gpurun_out/r2t_launches_step_${prec}_raw.csv \
      python tools/one_step.py --precision $prec > gpurun_out/r2t_ncu_step_${prec}.log 2>&1
  tail -1 gpurun_out/r2t_ncu_step_${prec}.log | cut -c1-200
  python tools/summarize_launches.py gpurun_out/r2t_launches_step_${prec}_raw.csv gpurun_out/r2t_launches_step_${prec}.csv gpurun_out/r2t_traffic_${prec}.json UnrealEgo 256 $prec \
      --title "ncu launch list of one forward step, UnrealEgo, batch 256, $prec (tools/gpu_job_r2t.sh)"
  head -14 gpurun_out/r2t_launches_step_${prec}.csv
done
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:gemm_tc --launch-skip 1 -c 4 \
    -o gpurun_out/r2t_gemm_bf16 python tools/one_step.py --precision bf16 > gpurun_out/r2t_ncu_gemm.log 2>&1
tail -2 gpurun_out/r2t_ncu_gemm.log | cut -c1-200
for b in 1 8 16 32 128; do
  timeout 300 python bench.py --batch $b --steps 50 --warmup 5 --only-headline 2>&1 | tail -1 | cut -c1-220
  timeout 300 python bench.py --batch $b --steps 50 --warmup 5 --only-headline --graph 2>&1 | tail -1 | cut -c1-220
done
timeout 600 python bench.py --workload train --precision bf16 --batch 256 --steps 10 --warmup 3 --dump gpurun_out/r2t_train_b256.json 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench.py --workload train --precision bf16 --batch 32 --steps 10 --warmup 3 --dump gpurun_out/r2t_train_b32.json 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench.py --workload e2e_rgb --batch 512 --steps 10 --warmup 3 --dump gpurun_out/r2t_e2e_rgb.json 2>&1 | tail -1 | cut -c1-300
