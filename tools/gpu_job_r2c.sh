#!/bin/bash
# Round-2 job 1 (first thing this round): the whole strict GPU suite + A/B timing of everything that had never been timed.
#   gpurun --timeout 2700 -- 'bash tools/gpu_job_r2c.sh > gpurun_out/r2c.log 2>&1'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
# 1. strict suite (no quarantine any more), no -x: every failure is wanted
timeout 1500 python -m pytest tests -m gpu -q -rfEs 2>&1 | tail -60 > gpurun_out/r2c_pytest.log; tail -25 gpurun_out/r2c_pytest.log
# 2. attention op alone, v1 vs wide
for prec in bf16x3 bf16; do
  for v in v1 wide; do
    if [ $v = wide ]; then export EGOTAP_ATTN=wide; else unset EGOTAP_ATTN; fi
    echo "variant $v"; timeout 300 python tools/attn_only.py 256 $prec 2>&1 | tail -1
  done
done
unset EGOTAP_ATTN
# 3. step-level A/B
for prec in bf16x3 bf16; do
  timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --dump gpurun_out/r2c_bench_${prec}_v1.json 2>&1 | tail -1 | cut -c1-400
  EGOTAP_ATTN=wide timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --dump gpurun_out/r2c_bench_${prec}_wide.json 2>&1 | tail -1 | cut -c1-400
  EGOTAP_EPI=coalesced timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --dump gpurun_out/r2c_bench_${prec}_coal.json 2>&1 | tail -1 | cut -c1-400
  EGOTAP_EPI=coalesced EGOTAP_ATTN=wide timeout 600 python bench.py --precision $prec --steps 20 --warmup 3 --dump gpurun_out/r2c_bench_${prec}_both.json 2>&1 | tail -1 | cut -c1-400
done
for f in gpurun_out/r2c_bench_*.json; do echo $f; python tools/summarize_bench.py $f | head -14; done
# 4. small batches: launched vs graph, split-K on / off
for b in 1 8 16 32 128; do
  timeout 300 python bench.py --batch $b --steps 50 --warmup 5 2>&1 | tail -1 | cut -c1-200
  EGOTAP_SPLITK=1 timeout 300 python bench.py --batch $b --steps 50 --warmup 5 2>&1 | tail -1 | cut -c1-200
  timeout 300 python bench.py --batch $b --steps 50 --warmup 5 --graph 2>&1 | tail -1 | cut -c1-200
  EGOTAP_SPLITK=1 timeout 300 python bench.py --batch $b --steps 50 --warmup 5 --graph 2>&1 | tail -1 | cut -c1-200
done
# 5. training step: first numbers
for cfg in "bf16 32" "bf16 256" "bf16x3 32"; do
  set -- $cfg
  timeout 600 python bench.py --workload train --precision $1 --batch $2 --steps 10 --warmup 3 \
      --dump gpurun_out/r2c_train_$1_b$2.json 2>&1 | tail -1 | cut -c1-1800
done
timeout 600 python bench.py --workload train --precision bf16 --batch 32 --steps 10 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench.py --workload train --precision bf16 --batch 32 --steps 10 --warmup 3 --persistent-bptt 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench.py --workload train --precision bf16 --batch 32 --steps 10 --warmup 3 --persistent-bptt --graph 2>&1 | tail -1 | cut -c1-300
timeout 600 python bench.py --workload lifting_gt --steps 10 --warmup 3 --dump gpurun_out/r2c_lifting_gt.json 2>&1 | tail -1 | cut -c1-600
# 6. ncu: wide attention full capture, launch list of one training step
EGOTAP_ATTN=wide timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_wide -c 1 \
    -o gpurun_out/r2c_attention_wide python tools/attn_only.py 64 bf16x3 > gpurun_out/r2c_ncu_attn.log 2>&1
tail -2 gpurun_out/r2c_ncu_attn.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r2c_launches_train.csv python bench.py --workload train --precision bf16 --batch 32 --steps 1 --warmup 3 \
    > gpurun_out/r2c_ncu_train.log 2>&1
tail -2 gpurun_out/r2c_ncu_train.log | cut -c1-300
