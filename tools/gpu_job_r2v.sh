#!/bin/bash
# Round-2 job: GPU suite after the LayerNorm-fold plumbing (opt-in), and an interleaved A/B of the default path against the
# previous commit's library (tools/libegotap_b200_base.so) to show the added epilogue parameters cost the default path nothing.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfEs 2>&1 | tail -8
for rep in 1 2 3; do
  for prec in bf16x3 bf16; do
    for lib in base cur; do
      if [ $lib = cur ]; then unset EGOTAP_B200_LIB; else export EGOTAP_B200_LIB=$PWD/tools/libegotap_b200_$lib.so; fi
      timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --only-headline --dump gpurun_out/r2v_${prec}_${lib}_$rep.json > /dev/null 2>&1
      python tools/summarize_bench.py gpurun_out/r2v_${prec}_${lib}_$rep.json 2>/dev/null | head -1 | cut -c1-200 | sed "s/^/$lib $rep /"
    done
  done
done
unset EGOTAP_B200_LIB
for lib in base cur; do python tools/summarize_bench.py gpurun_out/r2v_bf16_${lib}_3.json 2>/dev/null | head -9 | sed "s/^/$lib /"; done
