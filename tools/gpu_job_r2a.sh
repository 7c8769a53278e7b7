#!/bin/bash
# First hardware run of the training step + GT-heatmap kernel (written in round 1 after the GPU budget was spent):
#   gpurun --timeout 2400 -- 'bash tools/gpu_job_r2a.sh > gpurun_out/r2a.log 2>&1'
# 1. op-level parity of every training kernel and of gt_heatmap_kernel against their oracles
# 2. engine-level parity (gradients vs autograd oracle, determinism, module loop, CUDA-graph step)
# 3. first bench lines of config 5 (bf16 and parity mode, batch 32 = the reference's, and 256) with per-kernel tables
# 4. launch list of one training step under ncu (shares only)
set -x
export EGOTAP_STRICT_UNVERIFIED=1      # the quarantined first-hardware-run tests (tests/conftest.py) run as ordinary tests here
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 900 python -m pytest tests/test_train_kernels.py tests/test_gt_heatmaps.py -m gpu -q -x 2>&1 | tail -15
timeout 1200 python -m pytest tests/test_zz_train_gpu.py -m gpu -q 2>&1 | tail -25
for cfg in "bf16 32" "bf16 256" "bf16x3 32"; do
  set -- $cfg
  timeout 600 python bench.py --workload train --precision $1 --batch $2 --steps 10 --warmup 3 \
      --dump gpurun_out/r2a_train_$1_b$2.json 2>&1 | tail -1 | cut -c1-1500
done
timeout 600 python bench.py --workload train --precision bf16 --batch 32 --steps 10 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-600
timeout 600 python bench.py --workload train --precision bf16 --batch 32 --steps 10 --warmup 3 --persistent-bptt 2>&1 | tail -1 | cut -c1-600
timeout 600 python bench.py --workload train --precision bf16 --batch 32 --steps 10 --warmup 3 --persistent-bptt --graph 2>&1 | tail -1 | cut -c1-600
timeout 600 python bench.py --workload lifting_gt --steps 10 --warmup 3 --dump gpurun_out/r2a_lifting_gt.json 2>&1 | tail -1 | cut -c1-900
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r2a_launches_train.csv python bench.py --workload train --precision bf16 --batch 32 --steps 1 --warmup 3 \
    > gpurun_out/r2a_ncu_train.log 2>&1
tail -3 gpurun_out/r2a_ncu_train.log | cut -c1-300
