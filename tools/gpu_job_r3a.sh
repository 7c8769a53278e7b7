#!/bin/bash
# Round-2 job: L2 prefetch of the next item's K / V tiles in the attention producer (base = the same build without it):
# op-level A/B with bit comparison, then the headline step interleaved.
set -x
mkdir -p gpurun_out
for prec in bf16x3 bf16; do
  EGOTAP_B200_LIB=$PWD/tools/libegotap_b200_base.so timeout 120 python tools/attn_check.py 256 $prec 5 2>&1 | tail -1 | cut -c1-200 | sed "s/^/base /"
  timeout 120 python tools/attn_check.py 256 $prec 5 2>&1 | tail -1 | cut -c1-200 | sed "s/^/cur  /"
done
for rep in 1 2 3; do
for prec in bf16x3 bf16; do
  EGOTAP_B200_LIB=$PWD/tools/libegotap_b200_base.so timeout 300 python tools/attn_only.py 256 $prec 2>&1 | tail -1 | sed 's/^/base  /'
  timeout 300 python tools/attn_only.py 256 $prec 2>&1 | tail -1 | sed 's/^/new   /'
done; done
for rep in 1 2; do
  for prec in bf16x3 bf16; do
    for lib in base cur; do
      if [ $lib = cur ]; then unset EGOTAP_B200_LIB; else export EGOTAP_B200_LIB=$PWD/tools/libegotap_b200_$lib.so; fi
      timeout 600 python bench.py --precision $prec --steps 30 --warmup 3 --only-headline --dump gpurun_out/r3a_${prec}_${lib}_$rep.json > /dev/null 2>&1
      python tools/summarize_bench.py gpurun_out/r3a_${prec}_${lib}_$rep.json 2>/dev/null | grep -E "^step|attention" | cut -c1-170 | sed "s/^/$lib $rep /"
    done
  done
done
