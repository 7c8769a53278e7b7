#!/bin/bash
# Round-2 job: LayerNorm fold (the six in-layer LayerNorm passes folded into the GEMMs around them): GPU suite, interleaved A/B
# against EGOTAP_LN=kernel on the headline step in both precisions, per-kernel tables.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rfEs 2>&1 | tail -12
cut -c1-330 gpurun_out/parity.jsonl | head -12
for rep in 1 2 3; do
  for prec in bf16x3 bf16; do
    for ln in kernel fold; do
      EGOTAP_LN=$ln timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --only-headline --dump gpurun_out/r2u_${prec}_${ln}_$rep.json > /dev/null 2>&1
      python tools/summarize_bench.py gpurun_out/r2u_${prec}_${ln}_$rep.json 2>/dev/null | head -1 | cut -c1-200 | sed "s/^/$ln $rep /"
    done
  done
done
for ln in kernel fold; do for prec in bf16x3 bf16; do
  python tools/summarize_bench.py gpurun_out/r2u_${prec}_${ln}_3.json 2>/dev/null | head -16 | sed "s/^/$ln /"
done; done
