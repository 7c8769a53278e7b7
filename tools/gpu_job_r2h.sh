#!/bin/bash
# Round-2 multi-GPU job (run with gpurun --gpus N): the driver's line at N GPUs (headline + modes.bf16 + config3, one job-level
# gather), the training step with the staged gradient all-reduce, config 4.
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/r2h_bench_default_n$N.json 2> gpurun_out/r2h_bench_default_n$N.err; tail -3 gpurun_out/r2h_bench_default_n$N.err | cut -c1-300; cut -c1-400 gpurun_out/r2h_bench_default_n$N.json
for b in 32 256; do
  timeout 900 $TR bench.py --gpus $N --workload train --precision bf16 --batch $b --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2h_train_b${b}_n$N.json; cut -c1-300 gpurun_out/r2h_train_b${b}_n$N.json
done
timeout 900 $TR bench.py --gpus $N --workload e2e_rgb --batch 512 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2h_e2e_rgb_n$N.json; cut -c1-300 gpurun_out/r2h_e2e_rgb_n$N.json
