#!/bin/bash
# Round-2 final multi-GPU line (run with gpurun --gpus N): the driver's default line at N GPUs.
N=${1:-8}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/r2x_bench_default_n$N.json 2> gpurun_out/r2x_bench_default_n$N.err; tail -2 gpurun_out/r2x_bench_default_n$N.err | cut -c1-300
python - <<PY
import json
for l in open('gpurun_out/r2x_bench_default_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'rank_ms', d.get('rank_ms',{}).get('min'), d.get('rank_ms',{}).get('max'), 'bf16', d['modes']['bf16']['value'], d['modes']['bf16']['ms_per_step'], 'config3', d['config3']['value'], d['config3'].get('ms_per_step'))
PY
