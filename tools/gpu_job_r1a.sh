#!/bin/bash
# round-1 GPU job A: parity tests, first bench lines, ncu launch list, one full capture of the top GEMM
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --dump gpurun_out/bench_x3.json 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --precision bf16 --dump gpurun_out/bench_bf16.json 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1
# launch list of one step (169+ launches after the warm-up step's launches are skipped by -s)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv \
    --log-file gpurun_out/launches_x3.csv python tools/one_step.py --batch 256 > gpurun_out/launches_x3.log 2>&1
tail -2 gpurun_out/launches_x3.log
# full capture of the dominant GEMM launches (MLP-up / MLP-down / QKV shapes) at a smaller batch to bound replay time
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --nvtx --nvtx-include "timed/" \
    -s 4 -c 5 -o gpurun_out/gemm_full python tools/one_step.py --batch 128 > gpurun_out/gemm_full.log 2>&1
tail -2 gpurun_out/gemm_full.log
ls -la gpurun_out
