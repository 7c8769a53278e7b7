#!/bin/bash
# Round-2 job: LayerNorm fold vs LayerNorm kernels over LONG timed regions (settled power governor), interleaved.
set -x
mkdir -p gpurun_out
for rep in 1 2; do
  for cfg in "bf16x3 120" "bf16 300"; do
    set -- $cfg
    for ln in kernel fold; do
      EGOTAP_LN=$ln timeout 600 python bench.py --precision $1 --steps $2 --warmup 3 --only-headline --dump gpurun_out/r3b_$1_${ln}_$rep.json > /dev/null 2>&1
      python tools/summarize_bench.py gpurun_out/r3b_$1_${ln}_$rep.json 2>/dev/null | head -1 | cut -c1-230 | sed "s/^/$ln $rep /"
    done
  done
done
