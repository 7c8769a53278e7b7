#!/bin/bash
# Round-2 job: LayerNorm fold at the reference's evaluation batch sizes (not power-bound there), launched and CUDA-graph replay.
set -x
mkdir -p gpurun_out
for b in 1 16 32; do
  for ln in kernel fold; do
    for g in "" "--graph"; do
      EGOTAP_LN=$ln timeout 300 python bench.py --batch $b --steps 100 --warmup 5 --only-headline $g 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('batch $b $ln $g: %.4f ms  %.0f frames/s  parity rel %.2e' % (d['ms_per_step'], d['value'], d['parity']['rel']))"
    done
  done
done
