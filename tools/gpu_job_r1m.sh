#!/bin/bash
# 8-GPU scaling check: reference arm + own arm at N=8 (and N=4), as the driver launches them
set -x
nvidia-smi --query-gpu=index,name --format=csv | head -10
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-900
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --preset EgoCap --batch 128 --precision bf16 2>&1 | tail -1 | cut -c1-600
