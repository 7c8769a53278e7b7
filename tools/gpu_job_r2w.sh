#!/bin/bash
# Round-2 job: board power per kernel family (tools/power_probe.py) and the default bench line with the power record.
set -x
mkdir -p gpurun_out
timeout 600 python tools/power_probe.py 2.5 2>&1 | tail -12 | tee gpurun_out/r2w_power_probe.txt
timeout 900 python bench.py > gpurun_out/r2w_bench_default.json 2> gpurun_out/r2w_bench_default.err; tail -2 gpurun_out/r2w_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2w_bench_default.json'))
print(d['value'], d['e2e']['value'], d['clocks'], d['modes']['bf16']['value'], d['modes']['bf16'].get('clocks'), d['config3']['value'])
PY
