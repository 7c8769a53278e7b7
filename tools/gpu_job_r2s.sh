#!/bin/bash
# Round-2 job: TMA epilogue for fp32 outputs (gemm.cuh TEPI): GPU suite, interleaved A/B against the load / store epilogue
# (EGOTAP_EPI_TMA=0) on the headline step in both precisions and on the training step.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rfEs 2>&1 | tail -8
for rep in 1 2; do
  for prec in bf16 bf16x3; do
    for tma in 0 1; do
      EGOTAP_EPI_TMA=$tma timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --only-headline --dump gpurun_out/r2s_${prec}_tma${tma}_$rep.json > /dev/null 2>&1
      python tools/summarize_bench.py gpurun_out/r2s_${prec}_tma${tma}_$rep.json 2>/dev/null | head -1 | cut -c1-200 | sed "s/^/tma$tma $rep /"
    done
  done
done
for tma in 0 1; do for prec in bf16 bf16x3; do
  python tools/summarize_bench.py gpurun_out/r2s_${prec}_tma${tma}_2.json 2>/dev/null | grep -E "1024 1024 1|1024 256 1|4096 1024 1 |1024 4096 |attention|2048 16384|2048 8192" | sed "s/^/tma$tma /"
done; done
for tma in 0 1; do
  EGOTAP_EPI_TMA=$tma timeout 600 python bench.py --workload train --precision bf16 --batch 256 --steps 10 --warmup 3 --dump gpurun_out/r2s_train_b256_tma$tma.json 2>&1 | tail -1 | cut -c1-400
  EGOTAP_EPI_TMA=$tma timeout 600 python bench.py --workload train --precision bf16 --batch 32 --steps 10 --warmup 3 --dump gpurun_out/r2s_train_b32_tma$tma.json 2>&1 | tail -1 | cut -c1-400
done
