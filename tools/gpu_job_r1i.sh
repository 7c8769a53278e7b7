#!/bin/bash
# final round-1 evidence: launch list with DRAM traffic, full captures of the dominant kernels
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "timed/" --csv \
    --log-file gpurun_out/launches_final.csv python tools/one_step.py --batch 256 > gpurun_out/launches_final.log 2>&1
tail -1 gpurun_out/launches_final.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|attention|pu_chain|layernorm" --nvtx --nvtx-include "timed/" \
    -s 9 -c 8 -o gpurun_out/final_full python tools/one_step.py --batch 256 > gpurun_out/final_full.log 2>&1
tail -1 gpurun_out/final_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pu_chain" --nvtx --nvtx-include "timed/" \
    -c 1 -o gpurun_out/pu_full python tools/one_step.py --batch 256 > gpurun_out/pu_full.log 2>&1
tail -1 gpurun_out/pu_full.log
ls -la gpurun_out | tail -8
