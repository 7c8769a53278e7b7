#!/bin/bash
# Round-2 job: bisect the parity-mode failure of the two-issuer attention kernel: base (single issuer), v1 (two issuers, no register
# split), v2 (two issuers, register split, P handed over in one step), current.
set -x
mkdir -p gpurun_out
for lib in base v1 v2 cur; do
  for prec in bf16x3 bf16; do
    for B in 8 64 256; do
      if [ $lib = cur ]; then unset EGOTAP_B200_LIB; else export EGOTAP_B200_LIB=$PWD/tools/libegotap_b200_$lib.so; fi
      timeout 120 python tools/attn_check.py $B $prec 30 2>&1 | tail -4 | cut -c1-400 | sed "s/^/$lib /"
    done
  done
done
unset EGOTAP_B200_LIB
timeout 300 compute-sanitizer --tool memcheck python tools/attn_check.py 8 bf16x3 3 2>&1 | tail -30 | cut -c1-300
