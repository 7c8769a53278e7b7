#!/bin/bash
# Round-2 final multi-GPU sanity (run with gpurun --gpus 2): smoke(), the driver's line at N = 2 (headline + modes.bf16 + config3),
# the reference arm under torchrun, the training step at N = 2.
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/r2x_bench_default_n$N.json 2> gpurun_out/r2x_bench_default_n$N.err; tail -2 gpurun_out/r2x_bench_default_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/r2x_bench_default_n$N.json'))
print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'rank_ms', d.get('rank_ms'), 'bf16', d['modes']['bf16']['value'], 'config3', d['config3']['value'], d['config3'].get('ms_per_step'))
PY
timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
timeout 900 $TR bench.py --gpus $N --workload train --precision bf16 --batch 256 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2x_train_b256_n$N.json; cut -c1-300 gpurun_out/r2x_train_b256_n$N.json
