#!/bin/bash
# Round-2 job: (i) wait-cycle trace of the two-issuer attention kernel; (ii) interleaved A/B of three builds on the headline
# step: base (HEAD), cur (two-issuer attention + leaner GELU + L2 prefetch of the residual rows), v3 (cur without the prefetch).
set -x
mkdir -p gpurun_out
timeout 300 python tools/attn_trace.py 2>&1 | tail -12
for rep in 1 2 3; do
  for prec in bf16 bf16x3; do
    for lib in base cur v3; do
      if [ $lib = cur ]; then unset EGOTAP_B200_LIB; else export EGOTAP_B200_LIB=$PWD/tools/libegotap_b200_$lib.so; fi
      timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --only-headline --dump gpurun_out/r2r_${prec}_${lib}_$rep.json > /dev/null 2>&1
      python tools/summarize_bench.py gpurun_out/r2r_${prec}_${lib}_$rep.json 2>/dev/null | head -1 | cut -c1-200 | sed "s/^/$lib $rep /"
    done
  done
done
unset EGOTAP_B200_LIB
for lib in base cur v3; do for prec in bf16 bf16x3; do
  python tools/summarize_bench.py gpurun_out/r2r_${prec}_${lib}_3.json 2>/dev/null | grep -E "1024 1024 1|1024 256 1|4096 1024 1 |1024 4096 1 |attention" | sed "s/^/$lib /"
done; done
