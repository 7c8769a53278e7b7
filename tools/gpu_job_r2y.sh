#!/bin/bash
# Round-2 job: the headline step over LONG timed regions (several seconds), i.e. under the settled power governor.
set -x
mkdir -p gpurun_out
for cfg in "bf16x3 30" "bf16x3 150" "bf16 30" "bf16 400"; do
  set -- $cfg
  timeout 600 python bench.py --precision $1 --steps $2 --warmup 3 --only-headline --dump gpurun_out/r2y_$1_$2.json > /dev/null 2>&1
  python tools/summarize_bench.py gpurun_out/r2y_$1_$2.json 2>/dev/null | head -1 | cut -c1-260 | sed "s/^/steps $2: /"
done
